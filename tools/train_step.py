"""BASELINE configs[4]: DiMSUM-L/2 bf16 training step on synthetic latents, DDP over NCCL (one process per GPU).

    python tools/train_step.py --steps 5                         # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/train_step.py --steps 5

Mirrors the reference loop (dimsum/train.py:302-321: DDP wrap :180, AdamW(lr=1e-4, wd=0), clip_grad_norm_(1.0)) and the
GVP velocity loss (dimsum/transport/path.py:228-247 alpha=sin(pi t/2), sigma=cos(pi t/2); transport.py:137-146
loss = mean((model(xt, t, y) - ut)^2)).  The scan / conv / wavelet forward AND backward run on this repo's kernels;
the gradient all-reduce is DDP's bucketed NCCL all-reduce over NVLink, overlapped with backward.
Prints one JSON line: latents/s, ms per step, and the peak memory.
"""
import argparse
import json
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def gvp_plan(t, x0, x1):
    a, da = torch.sin(t * math.pi / 2), math.pi / 2 * torch.cos(t * math.pi / 2)
    s, ds = torch.cos(t * math.pi / 2), -math.pi / 2 * torch.sin(t * math.pi / 2)
    e = lambda v: v.view(-1, 1, 1, 1)
    return e(a) * x1 + e(s) * x0, e(da) * x1 + e(ds) * x0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--batch", type=int, default=32, help="per-GPU batch (SURVEY.md 8d config 5)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--model", default="DiM-L/2")
    ap.add_argument("--depth", type=int, default=None, help="override depth (smoke runs)")
    ap.add_argument("--graph", action="store_true",
                    help="capture forward + backward + clip + AdamW of one step in a CUDA graph and replay it (single GPU): the "
                         "step is ~3000 launches for 86 ms of device work, so eager execution is host-bound")
    ap.add_argument("--profile", default=None, help="write a torch.profiler kernel table of one extra step to this file")
    args = ap.parse_args()
    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("LOCAL_RANK", 0), ("WORLD_SIZE", 1)))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = True
    from dimsum_b200.models_dim import DiM, DiM_models
    torch.manual_seed(0)
    kw = dict(img_resolution=32, in_channels=4, num_classes=1000, label_dropout=0.1)
    with torch.device(dev):
        model = DiM_models[args.model](**kw) if args.depth is None else DiM(depth=args.depth, hidden_size=1024, **kw)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "adaLN_modulation" in n or n.startswith("final_layer.linear"):
                p.normal_(0, 0.02)
    model = model.to(dev).train()
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    use_graph = args.graph and world == 1
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0, capturable=use_graph, fused=True)   # same update, one kernel
    g = torch.Generator(device=dev).manual_seed(rank)
    amp = torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.dtype == "bf16")

    def draw():
        x1 = torch.randn(args.batch, 4, 32, 32, generator=g, device=dev)
        y = torch.randint(0, 1000, (args.batch,), generator=g, device=dev)
        t = torch.rand(args.batch, generator=g, device=dev)
        return x1, torch.randn(x1.shape, generator=g, device=dev), y, t

    def train_on(x1, x0, y, t):
        xt, ut = gvp_plan(t, x0, x1)
        with amp:
            out = ddp(xt, t, y)
        loss = (out.float() - ut).square().mean()
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step()
        return loss

    if use_graph:
        static = [b.clone() for b in draw()]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                       # warm-up on a side stream, as graph capture requires
            for _ in range(3):
                train_on(*static)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        with torch.cuda.graph(graph):
            static_loss = train_on(*static)

        def step():
            for dst, src in zip(static, draw()):
                dst.copy_(src)
            graph.replay()
            return static_loss
    else:
        def step():
            return train_on(*draw())

    for _ in range(args.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if args.profile and rank == 0:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            step()
            torch.cuda.synchronize()
        with open(args.profile, "w") as f:
            f.write(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=70, max_name_column_width=90))
    if rank == 0:
        print(json.dumps({"metric": "DiMSUM-L/2 train latents/s", "value": args.batch * world / (ms.item() * 1e-3),
                          "ms_per_step": ms.item(), "n_gpus": world, "per_gpu_batch": args.batch, "dtype": args.dtype,
                          "launch": "CUDA graph replay" if use_graph else "eager",
                          "loss": float(loss), "peak_mem_gb": torch.cuda.max_memory_allocated() / 2**30,
                          "missing_grads": [n for n, p in model.named_parameters() if p.grad is None and "cond_proj" not in n]}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
