"""BASELINE configs[4]: DiMSUM-L/2 bf16 training step on synthetic latents, DDP over NCCL (one process per GPU).

    python tools/train_step.py --steps 5                         # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py --steps 5
    python bench.py --workload train [--gpus N]                  # the same step behind the bench contract

Mirrors the reference loop (dimsum/train.py:302-321: DDP wrap :180, AdamW(lr=1e-4, wd=0), clip_grad_norm_(1.0)) and the
GVP velocity loss (dimsum/transport/path.py:228-247 alpha=sin(pi t/2), sigma=cos(pi t/2); transport.py:137-146
loss = mean((model(xt, t, y) - ut)^2)).  The scan / conv(+x_proj) / wavelet forward AND backward run on this repo's kernels;
the gradient all-reduce is DDP's bucketed NCCL all-reduce over NVLink, overlapped with backward.

The whole step -- forward, backward with the bucketed all-reduces, gradient clipping, fused AdamW -- is captured into ONE
CUDA graph and replayed (the step is ~3000 launches for ~60 ms of device work, so eager execution is host-bound).  With DDP
the capture follows PyTorch's recipe: NCCL async error handling off, DDP constructed and warmed up for 11 iterations on a
side stream, then the capture.  `--no-graph` launches eagerly.
"""
import argparse
import json
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gvp_plan(t, x0, x1):
    a, da = torch.sin(t * math.pi / 2), math.pi / 2 * torch.cos(t * math.pi / 2)
    s, ds = torch.cos(t * math.pi / 2), -math.pi / 2 * torch.sin(t * math.pi / 2)
    e = lambda v: v.view(-1, 1, 1, 1)
    return e(a) * x1 + e(s) * x0, e(da) * x1 + e(ds) * x0


class TrainStep:
    """One optimizer step of config 5 on this rank's GPU: `step()` draws a synthetic batch on the device and runs (or
    replays) forward + backward (+ DDP all-reduce) + clip + AdamW; returns the loss tensor."""

    def __init__(self, dev, rank, world, batch=32, dtype="bf16", model_name="DiM-L/2", depth=None, use_graph=True, res=32,
                 bucket_mb=100, shadows=True, fuse_clip=True):
        from dimsum_b200 import amp
        from dimsum_b200.models_dim import DiM, DiM_models
        self.dev, self.rank, self.world, self.batch, self.res = dev, rank, world, batch, res
        torch.manual_seed(0)
        kw = dict(img_resolution=res, in_channels=4, num_classes=1000, label_dropout=0.1,
                  ssm_cfg={"freeze_dead_cond_proj": True})     # DDP(find_unused_parameters=False) must not wait for dcond = None
        with torch.device(dev):
            model = DiM_models[model_name](**kw) if depth is None else DiM(depth=depth, hidden_size=1024, **kw)
        with torch.no_grad():
            for n, p in model.named_parameters():
                if "adaLN_modulation" in n or n.startswith("final_layer.linear"):
                    p.normal_(0, 0.02)
        self.model = model.to(dev).train()
        # bf16 copies of the GEMM weights, refreshed by one multi-tensor copy after the optimizer step, instead of autocast's
        # per-weight casts every forward and backward (dimsum_b200/amp.py); DIMSUM_BF16_SHADOWS=0 / shadows=False: plain autocast
        use_shadows = shadows and dtype == "bf16" and os.environ.get("DIMSUM_BF16_SHADOWS", "1") != "0"
        self.shadows = amp.Bf16Shadows(self.model) if use_shadows else None
        self.fuse_clip = fuse_clip and os.environ.get("DIMSUM_FUSED_CLIP", "1") != "0"
        self.amp = torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == "bf16")
        self.g = torch.Generator(device=dev).manual_seed(rank)
        self.use_graph = use_graph
        self.graph = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            # DDP is constructed on the side stream so that its bucket streams / events are captured consistently
            self.ddp = (torch.nn.parallel.DistributedDataParallel(self.model, device_ids=[dev.index], gradient_as_bucket_view=True,
                                                                  bucket_cap_mb=bucket_mb)
                        if world > 1 else self.model)
            self.opt = torch.optim.AdamW(self.model.parameters(), lr=1e-4, weight_decay=0, capturable=use_graph, fused=True)
            if use_graph:
                self.static = [b.clone() for b in self.draw()]
                for _ in range(11 if world > 1 else 3):        # DDP needs 11 warm-up iterations before capture
                    self.train_on(*self.static)
        torch.cuda.current_stream().wait_stream(side)
        if use_graph:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            self.graph = torch.cuda.CUDAGraph()
            self.opt.zero_grad(set_to_none=True)
            with torch.cuda.graph(self.graph):
                self.static_loss = self.train_on(*self.static)

    def draw(self):
        x1 = torch.randn(self.batch, 4, self.res, self.res, generator=self.g, device=self.dev)
        y = torch.randint(0, 1000, (self.batch,), generator=self.g, device=self.dev)
        t = torch.rand(self.batch, generator=self.g, device=self.dev)
        return x1, torch.randn(x1.shape, generator=self.g, device=self.dev), y, t

    def train_on(self, x1, x0, y, t):
        xt, ut = gvp_plan(t, x0, x1)
        with self.amp:
            out = self.ddp(xt, t, y)
        loss = (out.float() - ut).square().mean()
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        if self.fuse_clip:
            # clip_grad_norm_(1.0) = every gradient times min(1, 1 / (norm + 1e-6)).  The fused AdamW kernel divides the
            # gradients by `grad_scale` on the fly (the hook torch.amp.GradScaler uses) and stores them back, so the clip rides
            # on the optimizer's own pass instead of a separate read-modify-write of all gradients
            grads = [p.grad for p in self.model.parameters() if p.grad is not None]
            total = torch.nn.utils.get_total_norm(grads, foreach=True)
            self.opt.grad_scale = torch.clamp(total + 1e-6, min=1.0).float()
        else:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1.0)
        self.opt.step()
        if self.shadows is not None:
            self.shadows.refresh()
        return loss

    def step(self, batch=None):
        """`batch`: optional (x1, x0, y, t) device tensors (the end-to-end leg of bench.py copies them from pinned host memory)."""
        src = batch if batch is not None else self.draw()
        if self.graph is None:
            return self.train_on(*src)
        for dst, s in zip(self.static, src):
            dst.copy_(s, non_blocking=True)
        self.graph.replay()
        return self.static_loss

    def missing_grads(self):
        return [n for n, p in self.model.named_parameters() if p.requires_grad and p.grad is None and "cond_proj" not in n]


def init_dist():
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("TORCH_NCCL_ASYNC_ERROR_HANDLING", "0")     # required to capture NCCL collectives in a CUDA graph
        dist.init_process_group("nccl", device_id=dev)
    return rank, local, world, dev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (SURVEY.md 8d config 5)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--model", default="DiM-L/2")
    ap.add_argument("--depth", type=int, default=None, help="override depth (smoke runs)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of one extra step to this file")
    args = ap.parse_args()
    rank, local, world, dev = init_dist()
    torch.backends.cuda.matmul.allow_tf32 = True
    ts = TrainStep(dev, rank, world, args.batch, args.dtype, args.model, args.depth, use_graph=not args.no_graph)
    for _ in range(args.warmup):
        ts.step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = ts.step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True, with_stack=True) as prof:
            ts.step()
            torch.cuda.synchronize()
        with open(args.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=90))
        with open(args.profile + ".shapes", "w") as f:      # which call sites the framework glue (copies, cats, sums) comes from
            f.write(prof.key_averages(group_by_input_shape=True).table(sort_by="self_cuda_time_total", row_limit=80,
                                                                        max_name_column_width=60, max_shapes_column_width=90))
        with open(args.profile + ".stacks", "w") as f:
            f.write(prof.key_averages(group_by_stack_n=6).table(sort_by="self_cuda_time_total", row_limit=60,
                                                                 max_name_column_width=50, max_src_column_width=110))
    if rank == 0:
        print(json.dumps({"metric": "DiMSUM-L/2 train latents/s", "value": args.batch * world / (ms.item() * 1e-3),
                          "ms_per_step": ms.item(), "n_gpus": world, "per_gpu_batch": args.batch, "dtype": args.dtype,
                          "launch": "CUDA graph replay" if ts.graph is not None else "eager",
                          "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "missing_grads": ts.missing_grads()}))
    if world > 1:      # graph-captured NCCL communicators block in destroy_process_group: leave without the teardown
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
