"""Developer micro-benchmark of the individual kernels (CUDA events, L2 flushed between iterations).

    python tools/bench_ops.py [--json out.json]

Prints achieved algorithmic GB/s against MEASURED_PEAKS.json for scan fwd (inference + training), conv, wavelet.
Not the contract benchmark (that is bench.py); used to steer kernel work.
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peak():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def timeit(fn, iters=10, warmup=3, flush=None):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--json", default=None)
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--bwd", action="store_true", help="time the backward kernels at every shape (default: R <= 64)")
    args = ap.parse_args()
    from dimsum_b200 import causal_conv1d_cuda, selective_scan_cuda, wavelet_packet, scanning_orders as so
    pk = peak()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    rows = []
    N = 16
    shapes = [(256, 2048, 256), (64, 2048, 256), (8, 2048, 256), (512, 1024, 256), (128, 1024, 1024)]
    if args.quick:
        shapes = shapes[:1]
    for dtype in (torch.float32, torch.bfloat16):
        s = 4 if dtype == torch.float32 else 2
        for (R, D, L) in shapes:
            g = torch.Generator(device="cuda").manual_seed(0)
            xz = torch.randn(R, 2 * D, L, generator=g, device="cuda").to(dtype)
            u, z = xz[:, :D], xz[:, D:]
            delta = (0.5 * torch.rand(D, R, L, generator=g, device="cuda")).to(dtype).transpose(0, 1)
            A = -0.5 * torch.rand(D, N, generator=g, device="cuda")
            Bm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
            Cm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
            Dv = torch.randn(D, generator=g, device="cuda")
            bias = 0.5 * torch.rand(D, generator=g, device="cuda")
            w = torch.randn(D, 4, generator=g, device="cuda")
            cb = torch.randn(D, generator=g, device="cuda")
            # inference scan: reads u, delta, z, B, C; writes out_z
            by = s * (4 * R * D * L + 2 * R * N * L) + 4 * (D * N + 2 * D)
            med, best = timeit(lambda: selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, z, bias, True, need_out=False, need_x=False),
                               flush=flush)
            rows.append(dict(op="scan_fwd_infer", dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by / med / 1e6,
                             frac=by / med / 1e6 / pk))
            A_init = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(D, 1)      # S4D-real init form
            med, best = timeit(lambda: selective_scan_cuda.fwd(u, delta, A_init, Bm, Cm, Dv, z, bias, True, need_out=False,
                                                               need_x=False, a_arith=True), flush=flush)
            rows.append(dict(op="scan_fwd_infer_initA", dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by / med / 1e6,
                             frac=by / med / 1e6 / pk))
            by_t = by + s * R * D * L + 4 * R * D * ((L + 31) // 32 + 1) * 2 * N
            med, best = timeit(lambda: selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, z, bias, True), flush=flush)
            rows.append(dict(op="scan_fwd_train", dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by_t / med / 1e6,
                             frac=by_t / med / 1e6 / pk))
            by_c = 2 * s * R * D * L + 20 * D
            med, best = timeit(lambda: causal_conv1d_cuda.causal_conv1d_fwd(u, w, cb, True), flush=flush)
            rows.append(dict(op="conv_fwd", dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by_c / med / 1e6,
                             frac=by_c / med / 1e6 / pk))
            if R <= 64 or args.bwd:
                out, xck, out_z = selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, z, bias, True)
                dout = torch.randn(R, D, L, generator=g, device="cuda").to(dtype)
                by_b = s * (9 * R * D * L + 2 * R * N * L) + 4 * 2 * R * N * L + 4 * R * D * ((L + 31) // 32) * 2 * N
                med, best = timeit(lambda: selective_scan_cuda.bwd(u, delta, A, Bm, Cm, Dv, z, bias, dout, xck, out, None, True, True),
                                   flush=flush)
                rows.append(dict(op="scan_bwd", dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by_b / med / 1e6,
                                 frac=by_b / med / 1e6 / pk))
                by_cb = 3 * s * R * D * L
                med, best = timeit(lambda: causal_conv1d_cuda.causal_conv1d_bwd(u, w, cb, dout, None, True), flush=flush)
                rows.append(dict(op="conv_bwd", dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by_cb / med / 1e6,
                                 frac=by_cb / med / 1e6 / pk))
                del out, xck, out_z, dout
            del xz, u, z, delta, Bm, Cm
        # wavelet at the model shape: 512 rows, 16x16 tokens, 512 channels
        x = torch.randn(512, 256, 512, device="cuda").to(dtype)
        pos = so.as_index(so.reverse_permut_np(so.window_order(16, 4, False)), "cuda")
        by_w = 2 * s * x.numel()
        med, best = timeit(lambda: wavelet_packet(x, pos), flush=flush)
        rows.append(dict(op="wavelet_fwd", dtype=str(dtype), R=512, D=512, L=256, ms=med, ms_best=best, gbs=by_w / med / 1e6,
                         frac=by_w / med / 1e6 / pk))
        from dimsum_b200 import wavelet_packet_inverse
        med, best = timeit(lambda: wavelet_packet_inverse(x, pos), flush=flush)
        rows.append(dict(op="wavelet_inv", dtype=str(dtype), R=512, D=512, L=256, ms=med, ms_best=best, gbs=by_w / med / 1e6,
                         frac=by_w / med / 1e6 / pk))
    for r in rows:
        print(f"{r['op']:20s} {r['dtype']:15s} R={r['R']:4d} D={r['D']:5d} L={r['L']:5d}  {r['ms']:8.3f} ms (best {r['ms_best']:.3f})"
              f"  {r['gbs']:8.1f} GB/s  {100 * r['frac']:5.1f}% of measured peak {pk:.0f}")
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
