import sys, torch
sys.path.insert(0, ".")
from tools.bench_ops import timeit
from dimsum_b200 import selective_scan_cuda, causal_conv1d_cuda, scanning_orders as so
R, D, L, N = 256, 1024, 256, 16
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for dtype in (torch.float32, torch.bfloat16):
    s = 4 if dtype == torch.float32 else 2
    xz = torch.randn(R, 2 * D, L, generator=g, device="cuda").to(dtype)
    wc, cb = torch.randn(D, 4, generator=g, device="cuda"), torch.randn(D, generator=g, device="cuda")
    delta = (0.5 * torch.rand(D, R, L, generator=g, device="cuda")).to(dtype).transpose(0, 1)
    A = -0.5 * torch.rand(D, N, generator=g, device="cuda")
    Bm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype); Cm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
    Dv, bias = torch.randn(D, generator=g, device="cuda"), 0.5 * torch.rand(D, generator=g, device="cuda")
    by = s * (4 * R * D * L + 2 * R * N * L)
    for fam in ("sweep", "zigma", "jpeg"):
        for k in (0, 1, 3):
            perm = so.as_index(so.SCAN_ZOO[fam](16)[k], "cuda")
            u = causal_conv1d_cuda.causal_conv1d_fwd(xz[:, :D], wc, cb, True, perm=perm)
            mc, _ = timeit(lambda: causal_conv1d_cuda.causal_conv1d_fwd(xz[:, :D], wc, cb, True, perm=perm), flush=flush)
            ms, _ = timeit(lambda: selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, xz[:, D:], bias, True, need_out=False, need_x=False, perm=perm), flush=flush)
            print(f"{str(dtype):15s} {fam}[{k}]  conv {mc:.3f} ms ({100 * 2 * s * R * D * L / mc / 1e6 / 6463:.0f} %)   scan {ms:.3f} ms ({100 * by / ms / 1e6 / 6463:.0f} %)")
