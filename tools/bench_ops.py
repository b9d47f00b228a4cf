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
    ap.add_argument("--shapes", default=None, help="comma-separated RxDxL list replacing the default scan / conv shapes")
    ap.add_argument("--scan-only", action="store_true", help="only the scan / conv rows (skip wavelet, glue and gather kernels)")
    ap.add_argument("--attention", action="store_true", help="with --scan-only: also time the attention kernel against the library SDPA")
    ap.add_argument("--tag", default="", help="free text appended to every printed row (experiment label)")
    args = ap.parse_args()
    from dimsum_b200 import causal_conv1d_cuda, selective_scan_cuda, wavelet_packet, scanning_orders as so
    pk = peak()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
    rows = []
    N = 16
    shapes = [(256, 2048, 256), (64, 2048, 256), (8, 2048, 256), (512, 1024, 256), (128, 1024, 1024)]
    if args.quick:
        shapes = shapes[:1]
    if args.shapes:
        shapes = [tuple(int(v) for v in item.split("x")) for item in args.shapes.split(",")]
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
            if D % 64 == 0 and L % 8 == 0:
                xw = (torch.randn(64, D, generator=g, device="cuda") / D ** 0.5).to(dtype)
                by_x = 2 * s * R * D * L + s * R * 64 * L + s * 64 * D
                for tag, precise in (("conv_xproj", False),) + ((("conv_xproj_3xtf32", True),) if dtype == torch.float32 else ()):
                    med, best = timeit(lambda: causal_conv1d_cuda.conv_xproj_fwd(u, w, cb, xw, precise=precise, split=32), flush=flush)
                    rows.append(dict(op=tag, dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by_x / med / 1e6,
                                     frac=by_x / med / 1e6 / pk))
                # the two-step path it replaces: conv kernel + cuBLAS x_proj + the two B / C rearrange copies
                def two_step():
                    cu = causal_conv1d_cuda.causal_conv1d_fwd(u, w, cb, True)
                    xd = torch.bmm(cu.transpose(1, 2), xw.t().unsqueeze(0).expand(R, -1, -1)).reshape(R * L, -1)
                    Bc = xd[:, 32:48].view(R, L, 1, N).permute(0, 2, 3, 1).contiguous()
                    Cc = xd[:, 48:].view(R, L, 1, N).permute(0, 2, 3, 1).contiguous()
                    return cu, xd, Bc, Cc
                med, best = timeit(two_step, flush=flush)
                rows.append(dict(op="conv+xproj_2step", dtype=str(dtype), R=R, D=D, L=L, ms=med, ms_best=best, gbs=by_x / med / 1e6,
                                 frac=by_x / med / 1e6 / pk))
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
        if args.scan_only:
            continue
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
        # the order-carrying and glue kernels at the model's shapes (512 rows x 256 tokens, 512 / 1024 / 4096 channels)
        from dimsum_b200 import fused
        R, L, C = 512, 256, 1024
        g = torch.Generator(device="cuda").manual_seed(1)
        order = so.as_index(so.implicit_order(16, True, True), "cuda")
        inv = so.as_index(so.reverse_permut_np(so.implicit_order(16, True, True)), "cuda")
        h = torch.randn(R, L, C, generator=g, device="cuda")                      # fp32 residual stream
        xh = h[:, :, :C // 2].to(dtype) if dtype != torch.float32 else h[:, :, :C // 2]
        ada = (0.1 * torch.randn(R, 3 * C, generator=g, device="cuda")).to(dtype)
        sh, sc, gt = ada.chunk(3, dim=1)
        sh2, sc2, gt2 = (t[:, :C // 2] for t in (sh, sc, gt))
        m = torch.randn(R, L, C // 2, generator=g, device="cuda").to(dtype)

        def add(op, fn, nbytes, Cc):
            med, best = timeit(fn, flush=flush)
            rows.append(dict(op=op, dtype=str(dtype), R=R, D=Cc, L=L, ms=med, ms_best=best, gbs=nbytes / med / 1e6,
                             frac=nbytes / med / 1e6 / pk))

        n_half = R * L * (C // 2)
        add("modulate_order", lambda: fused.modulate(xh, sh2, sc2, order), 2 * s * n_half, C // 2)
        add("gate_residual_order", lambda: fused.gate_residual(xh, gt2, m, inv), 3 * s * n_half, C // 2)
        add("token_gather", lambda: so.token_gather(m, order), 2 * s * n_half, C // 2)
        xo = torch.randn(R, L, C, generator=g, device="cuda").to(dtype)
        w = torch.ones(C, device="cuda")
        add("add_rmsnorm", lambda: fused.add_rmsnorm(xo, h, w, 1e-5), (s + 4 + 4 + s) * R * L * C, C)
        add("norm_modulate", lambda: fused.norm_modulate(xo, h, w, 1e-5, sh, sc, want_residual=True), (s + 4 + 4 + s) * R * L * C, C)
        x12 = torch.randn(R * L, 8 * C, generator=g, device="cuda").to(dtype)
        add("gelu_mul", lambda: fused.gelu_mul(x12), 3 * s * R * L * 4 * C, 4 * C)
        del x12, xo, h, m
        # conv / scan read through a jpeg table (scan_type="jpeg_8" of the reference): the gather variants
        Rr, Dd = 256, 1024
        perm = so.as_index(so.SCAN_ZOO["jpeg"](16)[1], "cuda")
        xz = torch.randn(Rr, 2 * Dd, L, generator=g, device="cuda").to(dtype)
        wc, cb = torch.randn(Dd, 4, generator=g, device="cuda"), torch.randn(Dd, generator=g, device="cuda")
        med, best = timeit(lambda: causal_conv1d_cuda.causal_conv1d_fwd(xz[:, :Dd], wc, cb, True, perm=perm), flush=flush)
        by = 2 * s * Rr * Dd * L
        rows.append(dict(op="conv_fwd_perm", dtype=str(dtype), R=Rr, D=Dd, L=L, ms=med, ms_best=best, gbs=by / med / 1e6, frac=by / med / 1e6 / pk))
        u = causal_conv1d_cuda.causal_conv1d_fwd(xz[:, :Dd], wc, cb, True, perm=perm)
        delta = (0.5 * torch.rand(Dd, Rr, L, generator=g, device="cuda")).to(dtype).transpose(0, 1)
        A = -0.5 * torch.rand(Dd, N, generator=g, device="cuda")
        Bm = torch.randn(Rr, 1, N, L, generator=g, device="cuda").to(dtype)
        Cm = torch.randn(Rr, 1, N, L, generator=g, device="cuda").to(dtype)
        Dv, bias = torch.randn(Dd, generator=g, device="cuda"), 0.5 * torch.rand(Dd, generator=g, device="cuda")
        by = s * (4 * Rr * Dd * L + 2 * Rr * N * L)
        med, best = timeit(lambda: selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, xz[:, Dd:], bias, True, need_out=False,
                                                           need_x=False, perm=perm), flush=flush)
        rows.append(dict(op="scan_fwd_infer_perm", dtype=str(dtype), R=Rr, D=Dd, L=L, ms=med, ms_best=best, gbs=by / med / 1e6, frac=by / med / 1e6 / pk))
        del xz, u, delta, Bm, Cm
    if not args.scan_only or args.attention:
        # attention at the model's shapes: 512 CFG rows, 256 tokens, 8 (fusion block) and 16 (DiT block) heads of 64, fp32 / TF32
        from dimsum_b200.attention import attention
        import torch.nn.functional as F
        torch.backends.cuda.matmul.allow_tf32 = True
        for H, Nt in ((8, 256), (16, 256), (8, 1024)):
            Bt = 512
            qkv = torch.randn(Bt, Nt, 3, H, 64, device="cuda")
            q, k, v = qkv.permute(2, 0, 3, 1, 4).unbind(0)
            by_a = 4 * 4 * Bt * H * Nt * 64                        # q, k, v read + out written once
            fl = 4.0 * Bt * H * Nt * Nt * 64
            with torch.no_grad():
                med, best = timeit(lambda: attention(q, k, v), flush=flush)
                rows.append(dict(op=f"attention_h{H}_n{Nt}", dtype="torch.float32", R=Bt, D=H * 64, L=Nt, ms=med, ms_best=best,
                                 gbs=by_a / med / 1e6, frac=by_a / med / 1e6 / pk, tflops=fl / med / 1e9))
                med, best = timeit(lambda: F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(Bt, Nt, H * 64), flush=flush)
                rows.append(dict(op=f"sdpa_lib_h{H}_n{Nt}", dtype="torch.float32", R=Bt, D=H * 64, L=Nt, ms=med, ms_best=best,
                                 gbs=by_a / med / 1e6, frac=by_a / med / 1e6 / pk, tflops=fl / med / 1e9))
    for r in rows:
        print(f"{r['op']:20s} {r['dtype']:15s} R={r['R']:4d} D={r['D']:5d} L={r['L']:5d}  {r['ms']:8.3f} ms (best {r['ms_best']:.3f})"
              f"  {r['gbs']:8.1f} GB/s  {100 * r['frac']:5.1f}% of measured peak {pk:.0f} {args.tag}"
              + (f"  {r['tflops']:.1f} TFLOP/s" if "tflops" in r else ""))
    if args.json:
        json.dump(rows, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
