"""Contract benchmark: DiMSUM-L/2 forward latents/s (BASELINE.json metric) + selective-scan roofline.

    python bench.py --gpus 1 --steps 10 --warmup 3                     # this repo (hand-written sm_100a kernels)
    python bench.py --impl reference --gpus 1 --steps 1 --warmup 0     # reference CPU path (oracle port) on the host cores
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]): DiMSUM-L/2 256px class-conditional CFG sampling, 256 synthetic latents
(4x32x32 -> L=256 tokens) sharded contiguously over the N ranks; one STEP is one denoising evaluation of the 250-point
Euler grid: v = model.forward_with_cfg(x, t, y, cfg_scale=4) on 2*256/N rows and x += dt*v.  value = 256 latents / step
time (max over ranks) -- total work is fixed, so scaling is "strong".  Weights are random-init DiM-L/2 (459.9 M
parameters) with the adaLN-zero layers re-randomised (otherwise the network output is identically zero).
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOTAL_LATENTS = 256
CFG_SCALE = 4.0
NUM_GRID = 250
D_INNER, D_STATE, SEQ = 1024, 16, 256
RES = 32                         # latent side: 32 (256px, BASELINE configs[2]) or 64 (512px, configs[3]); --px sets RES and SEQ
CPU_SAMPLE_LATENTS = 8          # bounded CPU sample: batching helps the CPU path (0.18 -> 0.57 latents/s from 1 to 8 on 8 cores)
CPU_SAMPLE_CHOICES = (4, 8, 16)  # the reference arm tries these batch sizes once and times the best per-latent one


def build_model(device, res=None, seed=0):
    res = RES if res is None else res
    from dimsum_b200.models_dim import DiM_models
    torch.manual_seed(seed)
    with torch.device(device):
        model = DiM_models["DiM-L/2"](img_resolution=res, in_channels=4, num_classes=1000, label_dropout=0.1)
    g = torch.Generator(device="cpu").manual_seed(seed + 1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "adaLN_modulation" in n or n.startswith("final_layer.linear"):
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
    return model.to(device).eval()    # buffers built from numpy tables are created on the CPU


def make_inputs(n_total, res=None, seed=0):
    res = RES if res is None else res
    g = torch.Generator().manual_seed(seed)
    z = torch.randn(n_total, 4, res, res, generator=g)
    y = torch.randint(0, 1000, (n_total,), generator=g)
    return z, y


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        rows = [r[1:] for r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()
                and (self.t0 is None or self.t0 <= r[0] <= (self.t1 or r[0]) + 0.15)]
        window = "timed region"
        if not rows:   # region shorter than the sampling period: fall back to everything sampled while the process ran
            rows = [r[1:] for r in self.rows if len(r) >= 10 and r[2].replace(".", "").isdigit()]
            window = "whole run (timed region shorter than the 100 ms sampling period)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(float(r[1])) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(rows[0][2])), "reasons": reasons, "samples": len(sm),
                "window": window}


def measured_peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def measure_scan_traffic(rows, dtype):
    """DRAM bytes of ONE scan launch at this run's shape, measured now: a child process runs the launch under
    `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum` (tools/one_scan.py).  -> (bytes or None, how)."""
    import shutil
    ncu = shutil.which("ncu") or "/usr/local/cuda/bin/ncu"
    if not os.path.exists(ncu):
        return None, "ncu not found"
    cmd = [ncu, "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum", "--clock-control", "none", "-k", "regex:scan_fwd", "-c", "1",
           "--csv", sys.executable, os.path.join(ROOT, "tools", "one_scan.py"), str(rows), str(D_INNER), str(SEQ), dtype]
    try:
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=240).stdout
    except Exception as e:          # noqa: BLE001
        return None, "ncu run failed: %s" % type(e).__name__
    total, seen = 0.0, 0
    for line in out.splitlines():
        if "dram__bytes_" not in line:
            continue
        cols = [c.strip().strip('"') for c in line.split('","')]
        try:
            unit, val = cols[-2].lower(), float(cols[-1].replace(",", ""))
        except (ValueError, IndexError):
            continue
        total += val * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(unit, 1)
        seen += 1
    if seen != 2:
        return None, "could not parse the ncu output"
    return int(total), "measured in this run: ncu dram__bytes_read.sum + dram__bytes_write.sum of one launch at this shape (tools/one_scan.py)"


def scan_bytes(rows, s):
    """Algorithmic bytes of one inference scan launch: reads u, delta, z, B, C, A, D, dt_bias; writes y."""
    return s * (4 * rows * D_INNER * SEQ + 2 * rows * D_STATE * SEQ) + 4 * (D_INNER * D_STATE + 2 * D_INNER)


def workload_name(n_total):
    return ("DiMSUM-L/2 %dpx CFG denoising evaluation (BASELINE configs[%d]): %d latents sharded over the ranks, 2x rows with "
            "CFG, one Euler step of the 250-point grid per step" % (8 * RES, 2 if RES == 32 else 3, n_total))


def cpu_reference_step(sd, n_latents, seed=0):
    """One CFG denoising evaluation of the oracle port on the host cores; returns seconds."""
    from oracle import ref_model
    z, y = make_inputs(n_latents, seed=seed)
    x = torch.cat([z, z])
    yy = torch.cat([y, torch.full_like(y, 1000)])
    t = torch.full((2 * n_latents,), 0.5)
    t0 = time.perf_counter()
    with torch.no_grad():
        v = ref_model.dim_forward_with_cfg_oracle(sd, x, t, yy, CFG_SCALE)
    dt = time.perf_counter() - t0
    assert torch.isfinite(v).all()
    return dt


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle port of DiM.forward_with_cfg on selective_scan_ref /
    causal_conv1d_ref semantics) with all host threads, bounded sample of the same workload."""
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count())
    model = build_model("cpu")
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    # the per-latent throughput of the CPU path depends on the batch: try a few sizes once (this doubles as warm-up) and time
    # the best one, so the arm is not handicapped by an arbitrary sample size
    tried = {c: c / cpu_reference_step(sd, c) for c in CPU_SAMPLE_CHOICES}
    n = max(tried, key=tried.get)
    for _ in range(max(0, args.warmup - 1)):
        cpu_reference_step(sd, n)
    times = [cpu_reference_step(sd, n) for _ in range(max(1, args.steps))]
    sec = sum(times) / len(times)
    val = n / sec
    line = {
        "impl": "reference", "metric": "DiMSUM-L/2 fwd latents/s", "value": val, "unit": "latents/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.latents), "impl": "reference CPU path (oracle port), bounded sample per step",
                   "latents_per_step": n, "rows_per_step": 2 * n, "cfg_scale": CFG_SCALE, "tokens": SEQ, "px": 8 * RES,
                   "batch_sizes_tried_latents_per_s": {str(k): v for k, v in tried.items()}},
        "cpu_baseline": {"value": val, "unit": "latents/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{n} latents ({2 * n} CFG rows) x {len(times)} evaluation(s) of the 249-evaluation sampler"},
        "e2e": {"value": val, "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def train_cpu_sample(depth=1, latents=1):
    """Bounded CPU sample of the training step: forward + backward of `latents` latents through the oracle port of the
    reference CPU path, truncated to `depth` of the 16 blocks (the full step takes minutes on the host); returns seconds."""
    from oracle import ref_model
    model = build_model("cpu")
    sd = {k: v.detach().clone().requires_grad_(v.is_floating_point()) for k, v in model.state_dict().items()}
    z, y = make_inputs(latents)
    t = torch.full((latents,), 0.5)
    t0 = time.perf_counter()
    out = ref_model.dim_forward_oracle(sd, z, t, y, depth=depth, attn_every=0)
    out.square().mean().backward()
    return time.perf_counter() - t0


def run_train(args, rank, local_rank, world):
    """--workload train: BASELINE configs[4], DiMSUM-L/2 bf16 training step, DDP over NCCL, one CUDA graph per step."""
    import torch.distributed as dist
    from dimsum_b200 import _lib
    from tools.train_step import TrainStep, init_dist
    rank, local_rank, world, dev = init_dist()
    torch.backends.cuda.matmul.allow_tf32 = True
    B = args.train_batch
    ts = TrainStep(dev, rank, world, B, "bf16" if args.dtype != "fp32" else "fp32", use_graph=not args.no_graph,
                   bucket_mb=args.bucket_mb)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_step = None
    with ClockSampler(local_rank) as clocks:
        for _ in range(args.warmup):
            ts.step()
        barrier()
        l0 = _lib.launch_count()
        clocks.mark_start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            loss = ts.step()
        e1.record()
        barrier()
        clocks.mark_end()
    ms = e0.elapsed_time(e1) / args.steps
    assert torch.isfinite(loss).all()
    # end to end: the step's latents and labels come from pinned host memory, the loss goes back to the host
    x1_h = torch.randn(B, 4, RES, RES).pin_memory()
    y_h = torch.randint(0, 1000, (B,)).pin_memory()
    loss_h = torch.empty((), dtype=torch.float32).pin_memory()
    _, x0_d, _, t_d = ts.draw()
    for _ in range(2):
        ts.step((x1_h.to(dev, non_blocking=True), x0_d, y_h.to(dev, non_blocking=True), t_d))
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(args.steps):
        lv = ts.step((x1_h.to(dev, non_blocking=True), x0_d, y_h.to(dev, non_blocking=True), t_d))
        loss_h.copy_(lv.detach(), non_blocking=True)
    a1.record()
    barrier()
    ms_e2e = a0.elapsed_time(a1) / args.steps
    # the dominant kernel of this repo in the training step is the scan backward: time its launches in one eager step
    bwd_events, fwd_events = [], []
    real_call = _lib.call

    def timed_call(name, params, stream):
        if name not in ("dimsum_selective_scan_bwd", "dimsum_selective_scan_fwd"):
            return real_call(name, params, stream)
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s_.record()
        real_call(name, params, stream)
        e_.record()
        (bwd_events if name.endswith("bwd") else fwd_events).append((s_, e_))

    _lib.call = timed_call
    before = _lib.launch_count()
    ts.train_on(*ts.draw())
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - before
    _lib.call = real_call
    bwd_ms = sum(s_.elapsed_time(e_) for s_, e_ in bwd_events) / max(1, len(bwd_events))
    fwd_ms = sum(s_.elapsed_time(e_) for s_, e_ in fwd_events) / max(1, len(fwd_events))
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30
    if world > 1:
        tt = torch.tensor([ms, ms_e2e, bwd_ms, fwd_ms, peak_mem], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e, bwd_ms, fwd_ms, peak_mem = tt.tolist()
    if rank == 0:
        s_b = 2 if args.dtype != "fp32" else 4
        peak, how = measured_peak()
        L, D, N = SEQ, D_INNER, D_STATE
        by_b = s_b * (9 * B * D * L + 2 * B * N * L) + 4 * 2 * B * N * L + 4 * B * D * ((L + 31) // 32) * 2 * N
        line = {
            "metric": "DiMSUM-L/2 train latents/s", "value": B * world / (ms * 1e-3), "unit": "latents/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16" if args.dtype != "fp32" else "f32", "data": "synthetic",
            "config": {"workload": "DiMSUM-L/2 bf16 training step on synthetic latents (BASELINE configs[4]): %d latents per GPU, GVP "
                                   "velocity loss, DDP gradient all-reduce over NCCL, clip 1.0, fused AdamW lr 1e-4" % B,
                       "per_gpu_batch": B, "ddp_bucket_mb": args.bucket_mb, "tokens": SEQ, "d_inner": D_INNER, "d_state": D_STATE,
                       "launch": "eager" if ts.graph is None else "one CUDA graph per step (forward, backward with DDP's bucketed NCCL "
                                 "all-reduces, clip, AdamW)",
                       "l2": "activations of one step >> 126 MB L2", "params": sum(p.numel() for p in ts.model.parameters()),
                       "peak_mem_gb": peak_mem},
            "e2e": {"value": B * world / (ms_e2e * 1e-3), "unit": "latents/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": (x1_h.numel() * 4 + y_h.numel() * 8) * world, "d2h_bytes_per_step": 4 * world},
            "gpu_launches": launches_per_step * args.steps * world,
            "roofline": {"kernel": "scan_bwd_kernel (selective scan backward)", "bound": "hbm",
                         "achieved": by_b / (bwd_ms * 1e-3) / 1e9, "peak": peak,
                         "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs, burst copy)", "unit": "GB/s",
                         "frac": by_b / (bwd_ms * 1e-3) / 1e9 / peak, "traffic": None, "algorithmic_bytes_per_launch": by_b,
                         "avg_launch_ms": bwd_ms, "launches_timed": len(bwd_events),
                         "scan_fwd_train_avg_launch_ms": fwd_ms, "timed_in": "one instrumented eager step after the timed region",
                         "share_of_step": bwd_ms * len(bwd_events) / ms},
            "clocks": clocks.summary(),
        }
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count())
            sec = train_cpu_sample(depth=1, latents=1)
            line["cpu_baseline"] = {"value": 1.0 / (sec * 16), "unit": "latents/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": "forward + backward of 1 latent through 1 of the 16 DiM-L/2 blocks of the oracle port "
                                              "(autograd through selective_scan_ref / causal_conv1d_ref semantics), %.1f s, scaled x16; no "
                                              "optimizer step" % sec}
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators that were captured into a CUDA graph do not tear down cleanly (destroy_process_group blocks):
        # everything is printed and flushed, leave without the teardown
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


def run_b200(args, rank, local_rank, world):
    import torch.distributed as dist
    from dimsum_b200 import _lib
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True      # reference: dimsum/train.py:20-21
    torch.backends.cudnn.allow_tf32 = True
    os.environ["DIMSUM_SCAN_ARITH"] = "1" if args.init_form_fastpath else "0"
    model = build_model(dev)
    autocast = torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.dtype == "bf16")
    n_total = args.latents
    assert n_total % world == 0
    n = n_total // world
    z_all, y_all = make_inputs(n_total)
    lo = rank * n
    z, y = z_all[lo:lo + n], y_all[lo:lo + n]
    x_host = torch.cat([z, z]).pin_memory()
    y_host = torch.cat([y, torch.full_like(y, 1000)]).pin_memory()
    t_host = torch.full((2 * n,), 0.5).pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    x0 = x_host.to(dev)
    yy = y_host.to(dev)
    ts = torch.linspace(0, 1, NUM_GRID, device=dev)
    ones = torch.ones(2 * n, device=dev)

    scan_events = []
    real_call = _lib.call

    def timed_call(name, params, stream):
        if name != "dimsum_selective_scan_fwd" or not timed_call.on:
            return real_call(name, params, stream)
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        real_call(name, params, stream)
        e.record()
        scan_events.append((s, e))

    timed_call.on = False
    _lib.call = timed_call

    def eager_step(x, i):
        with torch.no_grad(), autocast:
            v = model.forward_with_cfg(x, ones * ts[i], yy, cfg_scale=CFG_SCALE)
        return x + (ts[i + 1] - ts[i]) * v.float()

    graphed = None
    if not args.no_graph:
        from dimsum_b200.sampler import GraphedCfgStep
        graphed = GraphedCfgStep(model, x0, yy, CFG_SCALE, torch.bfloat16 if args.dtype == "bf16" else None)

    def step(x, i):
        if graphed is None:
            return eager_step(x, i)
        return graphed.euler(x, ones * ts[i], ts[i + 1] - ts[i])     # forward + fused CFG combine + Euler update, one replay

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.ncu_range:
        x = x0.clone()
        for i in range(args.warmup):
            x = eager_step(x, i)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        x = eager_step(x, args.warmup)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        print("ncu range done (no bench line: profiler runs are not measurements)")
        return

    # ------------------------------------------------------------------ device-resident timing
    with ClockSampler(local_rank) as clocks:       # nvidia-smi is started before the warm-up so it is sampling by the time
        x = x0.clone()                             # the timed region starts; only samples inside the region are reported
        for i in range(args.warmup):
            x = step(x, i)
        barrier()
        launches0 = _lib.launch_count()
        timed_call.on = True
        barrier()
        clocks.mark_start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(args.steps):
            x = step(x, args.warmup + i)
        ev1.record()
        barrier()
        clocks.mark_end()
        timed_call.on = False
    launches = _lib.launch_count() - launches0
    if graphed is not None:
        launches = graphed.launches_per_replay * args.steps     # replays do not pass through the C-ABI counter
    ms = ev0.elapsed_time(ev1) / args.steps
    assert torch.isfinite(x).all()
    if graphed is not None:
        # per-kernel CUDA events cannot ride inside a graph replay: time the scan launches in an instrumented eager
        # pass over the same inputs, right after the timed region
        timed_call.on = True
        xe = x0.clone()
        for i in range(min(args.steps, 3)):
            xe = eager_step(xe, i)
        torch.cuda.synchronize()
        timed_call.on = False
    scan_ms = sum(s.elapsed_time(e) for s, e in scan_events) / max(1, len(scan_events))
    n_general = len(scan_events)
    # the same launches with the init-form (arithmetic-progression A) shortcut, for the record
    fast_ms = None
    if not args.init_form_fastpath:
        os.environ["DIMSUM_SCAN_ARITH"] = "1"
        for m in model.modules():
            if hasattr(m, "_arith_key"):
                m._arith_key = None
        eager_step(x0.clone(), 0)                      # host check + cache (one sync per mixer)
        torch.cuda.synchronize()
        timed_call.on = True
        xe = x0.clone()
        for i in range(2):
            xe = eager_step(xe, i)
        torch.cuda.synchronize()
        timed_call.on = False
        fast = scan_events[n_general:]
        fast_ms = sum(s.elapsed_time(e) for s, e in fast) / max(1, len(fast))
        os.environ["DIMSUM_SCAN_ARITH"] = "0"
        for m in model.modules():
            if hasattr(m, "_arith_key"):
                m._arith_key = None

    # the same scan launch alone after a short idle pause (the step loop leaves the GPU power-capped; the kernel is
    # issue/SFU-bound, so its time follows the SM clock), L2 flushed each time
    iso_ms = None
    try:
        time.sleep(3.0)
        from dimsum_b200 import selective_scan_cuda
        R_, dt_ = 2 * n, (torch.float32 if args.dtype == "fp32" else torch.bfloat16)
        gi = torch.Generator(device=dev).manual_seed(5)
        xz_i = torch.randn(2 * D_INNER, R_, SEQ, generator=gi, device=dev).to(dt_).transpose(0, 1)
        de_i = (0.5 * torch.rand(D_INNER, R_, SEQ, generator=gi, device=dev)).to(dt_).transpose(0, 1)
        u_i = torch.randn(R_, D_INNER, SEQ, generator=gi, device=dev).to(dt_)
        A_i = -0.5 * torch.rand(D_INNER, D_STATE, generator=gi, device=dev) - 0.05
        B_i = torch.randn(R_, 1, D_STATE, SEQ, generator=gi, device=dev).to(dt_)
        C_i = torch.randn(R_, 1, D_STATE, SEQ, generator=gi, device=dev).to(dt_)
        Dv_i, bi_i = torch.ones(D_INNER, device=dev), torch.rand(D_INNER, device=dev) - 3.0
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        run = lambda: selective_scan_cuda.fwd(u_i, de_i, A_i, B_i, C_i, Dv_i, xz_i[:, D_INNER:], bi_i, True, need_out=False, need_x=False)
        for _ in range(3):
            run()
        ts_i = []
        for _ in range(7):
            flush.zero_()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(); run(); a1.record()
            torch.cuda.synchronize()
            ts_i.append(a0.elapsed_time(a1))
        iso_ms = sorted(ts_i)[len(ts_i) // 2]
        del xz_i, de_i, u_i, B_i, C_i, flush
    except Exception:
        iso_ms = None

    # ------------------------------------------------------------------ end to end: host buffers in, host buffers out
    for i in range(min(2, args.warmup)):
        xd = x_host.to(dev, non_blocking=True)
        out_host.copy_(step(xd, i), non_blocking=True)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        xd = x_host.to(dev, non_blocking=True)
        yd = y_host.to(dev, non_blocking=True)
        td = t_host.to(dev, non_blocking=True)
        if graphed is not None:
            graphed.y.copy_(yd, non_blocking=True)
            out_host.copy_(graphed.euler(xd, td, 1.0 / (NUM_GRID - 1)), non_blocking=True)
        else:
            with torch.no_grad(), autocast:
                v = model.forward_with_cfg(xd, td, yd, cfg_scale=CFG_SCALE)
            out_host.copy_(xd + (1.0 / (NUM_GRID - 1)) * v.float(), non_blocking=True)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps

    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        xe = x0.clone()
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(2):
                xe = eager_step(xe, i)
            torch.cuda.synchronize()
        with open(args.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=100))

    gather_check = None
    if world > 1:
        # the one data-path collective of the sampling design: sample_cfg_sharded's NCCL all_gather_into_tensor of the final
        # latents (dimsum_b200/sampler.py).  Run it once here on hardware -- a short 3-point Euler grid over the same sharded
        # batch -- and check every rank receives exactly the slices the ranks computed locally.
        from dimsum_b200.sampler import sample_cfg, sample_cfg_sharded
        zs, ys = z_all[: 8 * world].to(dev), y_all[: 8 * world].to(dev)
        with autocast:
            full = sample_cfg_sharded(model, zs, ys, cfg_scale=CFG_SCALE, num_steps=3)
            mine = sample_cfg(model, zs[rank * 8:(rank + 1) * 8], ys[rank * 8:(rank + 1) * 8], cfg_scale=CFG_SCALE, num_steps=3)
        same_local = torch.equal(full[rank * 8:(rank + 1) * 8], mine)
        digest = full.double().sum().reshape(1)                    # identical on every rank iff every rank got every slice
        lo_d, hi_d = digest.clone(), digest.clone()
        dist.all_reduce(lo_d, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi_d, op=dist.ReduceOp.MAX)
        flag = torch.tensor([int(same_local and bool(torch.isfinite(full).all()) and lo_d.item() == hi_d.item())], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        gather_check = {"collective": "NCCL all_gather_into_tensor of the final latents (sampler.sample_cfg_sharded)",
                        "latents": 8 * world, "grid_points": 3, "bytes_gathered_per_rank": int(full.numel() * full.element_size()),
                        "every_rank_holds_every_rank_local_result": bool(flag.item())}
        assert flag.item() == 1, "sample_cfg_sharded: gathered latents differ from the rank-local results"
        tt = torch.tensor([ms, ms_e2e, scan_ms, fast_ms or 0.0], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms, ms_e2e, scan_ms, fast_ms = tt.tolist()
        fast_ms = fast_ms or None
        lc = torch.tensor([launches], device=dev)
        dist.all_reduce(lc)
        launches = int(lc.item())

    if rank == 0:
        s = 4 if args.dtype == "fp32" else 2
        peak, how = measured_peak()
        by = scan_bytes(2 * n, s)
        achieved = by / (scan_ms * 1e-3) / 1e9
        if world > 1:
            traffic, traffic_how = None, "not measured in multi-rank runs (ncu is single-process only); see the 1-GPU line"
        elif args.no_traffic:
            traffic, traffic_how = None, "skipped (--no-traffic)"
        else:
            traffic, traffic_how = measure_scan_traffic(2 * n, args.dtype)
        # the kernel's second ceiling: 20 MUFU results per element (16 decays + softplus 2 + silu 2) at the measured 16 per clock
        # per SM (profiles/r1f_pipe_microbench.md) against the bytes per element at the HBM peak
        elems = 2 * n * D_INNER * SEQ
        sm_clock = 1e6 * (clocks.summary().get("sm_mhz") or 1965)
        sfu_floor_ms = elems * 20 / 16 / 148 / sm_clock * 1e3
        sfu_floor_ms_max_clock = elems * 20 / 16 / 148 / 1.965e9 * 1e3
        line = {
            "metric": "DiMSUM-L/2 fwd latents/s", "value": n_total / (ms * 1e-3), "unit": "latents/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if args.dtype == "fp32" else "bf16", "data": "synthetic",
            "config": {"workload": workload_name(n_total),
                       "latents_total": n_total, "rows_per_rank": 2 * n, "tokens": SEQ, "d_inner": D_INNER, "d_state": D_STATE,
                       "cfg_scale": CFG_SCALE, "matmul": "tf32" if args.dtype == "fp32" else "bf16 autocast",
                       "launch": "eager" if graphed is None else "CUDA graph replay (%d launches of this repo's kernels per step)"
                                 % graphed.launches_per_replay,
                       "l2": "working set per step >> 126 MB L2 (xz alone is %.0f MB per mixer call)" % (2 * n * 2 * D_INNER * SEQ * s / 1e6),
                       "params": sum(p.numel() for p in model.parameters())},
            "e2e": {"value": n_total / (ms_e2e * 1e-3), "unit": "latents/s", "ms_per_step": ms_e2e,
                    "h2d_bytes_per_step": (x_host.numel() * 4 + y_host.numel() * 8 + t_host.numel() * 4) * world,
                    "d2h_bytes_per_step": out_host.numel() * 4 * world},
            "gpu_launches": launches,
            "roofline": {"kernel": "scan_fwd_kernel (selective scan forward, inference)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "peak_source": how + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_how,
                         "algorithmic_bytes_per_launch": by,
                         "sfu_ceiling": {"mufu_results_per_element": 20, "mufu_results_per_clk_per_sm": 16,
                                         "floor_ms_at_step_clock": sfu_floor_ms, "frac_of_sfu_ceiling_in_step": sfu_floor_ms / scan_ms,
                                         "floor_ms_at_max_clock": sfu_floor_ms_max_clock,
                                         "hbm_frac_at_sfu_ceiling": by / (sfu_floor_ms_max_clock * 1e-3) / 1e9 / peak,
                                         "note": "general A needs 20 MUFU results per element; the kernel is bound by the SFU pipe, "
                                                 "not by HBM (ncu: XU pipe 84.5 % busy, profiles/r1e_scan_fwd_ncu.md); `frac` is "
                                                 "against HBM as the contract asks, this block gives the fraction of the binding unit"},
                         "avg_launch_ms": scan_ms, "launches_timed": n_general,
                         "isolated": None if iso_ms is None else
                             {"avg_launch_ms": iso_ms, "frac": by / (iso_ms * 1e-3) / 1e9 / peak,
                              "note": "identical launch (same shapes and strides, general A, L2 flushed) timed alone after a 3 s idle "
                                      "pause: the gap to avg_launch_ms is the SM clock under sw_power_cap during the step loop, "
                                      "not interference (tools/layout_probe.py: placement of u/delta/z does not matter)"},
                         "path": "one exp per step (A rows arithmetic, --init-form-fastpath)" if args.init_form_fastpath
                                 else "general A (16 exps per step)",
                         "init_form_A_fastpath": None if fast_ms is None else
                             {"avg_launch_ms": fast_ms, "achieved": by / (fast_ms * 1e-3) / 1e9, "frac": by / (fast_ms * 1e-3) / 1e9 / peak,
                              "note": "same launches with the shortcut for A[d][n] = (n+1) A[d][0] (S4D-real init, true for this "
                                      "random-init model, not for trained checkpoints); not used for `value`"},
                         "timed_in": "timed region" if graphed is None else "instrumented eager pass right after the timed region",
                         "share_of_step": scan_ms * 32 / ms},
            "clocks": clocks.summary(),
        }
        if gather_check is not None:
            line["all_gather_check"] = gather_check
        if world == 1 and not args.no_cpu_baseline:
            torch.set_num_threads(os.cpu_count())
            sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
            sec = cpu_reference_step(sd, CPU_SAMPLE_LATENTS)
            line["cpu_baseline"] = {"value": CPU_SAMPLE_LATENTS / sec, "unit": "latents/s", "cores": torch.get_num_threads(),
                                    "kind": "port",
                                    "sample": "%d latents (%d CFG rows), 1 of the 249 evaluations, oracle port of the reference "
                                              "CPU path (selective_scan_ref / causal_conv1d_ref semantics), %.1f s"
                                              % (CPU_SAMPLE_LATENTS, 2 * CPU_SAMPLE_LATENTS, sec)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"])
    ap.add_argument("--workload", default="sample", choices=["sample", "train"],
                    help="sample: CFG denoising evaluation (BASELINE configs[2], the headline); train: bf16 DDP training step (configs[4])")
    ap.add_argument("--train-batch", type=int, default=32, help="per-GPU batch of --workload train (SURVEY.md 8d config 5)")
    ap.add_argument("--bucket-mb", type=int, default=100, help="DDP gradient bucket size of --workload train (MB)")
    ap.add_argument("--latents", type=int, default=TOTAL_LATENTS)
    ap.add_argument("--px", type=int, default=256, choices=[256, 512], help="image size: 256 (L=256 tokens) or 512 (L=1024, configs[3])")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-traffic", action="store_true", help="skip the in-run ncu measurement of the scan's DRAM traffic (~30 s)")
    ap.add_argument("--profile", default=None,
                    help="also write a torch.profiler kernel table of two eager steps to this file (after all timing)")
    ap.add_argument("--ncu-range", action="store_true",
                    help="for `ncu --profile-from-start off`: warm up, then bracket ONE eager step with cudaProfilerStart/Stop and exit "
                         "(launch lists under profiles/; numbers printed under a profiler are never bench values)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    ap.add_argument("--init-form-fastpath", action="store_true",
                    help="let the scan use its one-exp-per-step path for arithmetic-progression A.  The random-init benchmark "
                         "model satisfies it (A = -(1..16)), trained checkpoints do not, so the headline runs the GENERAL path")
    args = ap.parse_args()
    global RES, SEQ
    RES = args.px // 8
    SEQ = (RES // 2) ** 2
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if args.workload == "train":
        if args.dtype == "fp32" and "--dtype" not in sys.argv:
            args.dtype = "bf16"                       # configs[4] is a bf16-autocast config
        run_train(args, rank, local_rank, world)
        return
    run_b200(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
