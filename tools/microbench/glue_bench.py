"""Timing of the training-glue kernels on their DiM-L/2 shapes (CUDA events, 256 MB L2 flush between iterations):
    python tools/microbench/glue_bench.py
token_colsum (bias gradients, adaLN gradient reductions) next to torch.sum, add_rmsnorm backward, modulate / gate kernels."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from dimsum_b200 import fused  # noqa: E402

PEAK = 6462.7


def timeit(fn, iters=10):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(3):
        fn()
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def row(name, ms, nbytes):
    print(f"| {name} | {ms * 1e3:.1f} us | {nbytes / ms / 1e6:.0f} GB/s | {100 * nbytes / ms / 1e6 / PEAK:.1f} % |")


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    print("| kernel, shape | time | algorithmic bandwidth | of HBM peak |\n|---|---:|---:|---:|")
    for B, L, C in ((32, 256, 512), (32, 256, 1024), (32, 256, 1536), (32, 256, 8192)):
        gr = torch.randn(B, L, C, generator=g, device="cuda").bfloat16()
        x = torch.randn(B, L, C, generator=g, device="cuda")
        idx = torch.randperm(L, device="cuda").int()
        row(f"token_colsum g bf16 ({B},{L},{C})", timeit(lambda: fused.token_colsum(gr, out_dtype=torch.float32)), gr.numel() * 2)
        row(f"torch sum(0) of ({B * L},{C}) bf16", timeit(lambda: gr.view(-1, C).sum(0)), gr.numel() * 2)
        if C <= 1024:
            row(f"token_colsum g bf16, x fp32 ({B},{L},{C})", timeit(lambda: fused.token_colsum(gr, x, out_dtype=torch.float32)), gr.numel() * 6)
            row(f"token_colsum g bf16, x fp32 through a table ({B},{L},{C})",
                timeit(lambda: fused.token_colsum(gr, x, out_dtype=torch.float32, x_idx=idx)), gr.numel() * 6)
    B, L, C = 32, 256, 1024
    x = torch.randn(B, L, C, generator=g, device="cuda").bfloat16().requires_grad_(True)
    res = torch.randn(B, L, C, generator=g, device="cuda").requires_grad_(True)
    w = torch.ones(C, device="cuda", requires_grad=True)
    y, h = fused.add_rmsnorm_fn(x, res, w, 1e-5, out_dtype=torch.float32)
    gy, gh = torch.randn_like(y), torch.randn_like(h)
    n = B * L * C
    row("add_rmsnorm backward (dy fp32, dres fp32 -> dx bf16, dres fp32)",
        timeit(lambda: torch.autograd.grad((y, h), (x, res, w), (gy, gh), retain_graph=True)), n * (4 + 4 + 4 + 2 + 4))
    sh = torch.randn(B, 3 * C, generator=g, device="cuda").bfloat16()
    row("modulate fp32 -> bf16", timeit(lambda: fused.modulate(res.detach(), sh[:, :C], sh[:, C:2 * C], out_dtype=torch.bfloat16)), n * 6)
    row("gate_residual fp32 + bf16 -> fp32", timeit(lambda: fused.gate_residual(res.detach(), sh[:, :C], x.detach())), n * 10)


if __name__ == "__main__":
    main()
