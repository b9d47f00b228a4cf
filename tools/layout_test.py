"""Developer experiment: how does the scan time move with the number of MUFU ops per element?"""
import os, subprocess, sys
import torch
sys.path.insert(0, "/root/repo")
if len(sys.argv) == 1:
    for poly in ("0", "2"):
        env = dict(os.environ, DIMSUM_SCAN_POLY=poly)
        subprocess.run([sys.executable, __file__, "child"], env=env)
    sys.exit(0)
from dimsum_b200 import selective_scan_cuda
from tools.bench_ops import timeit
R, D, L, N = 512, 1024, 256, 16
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for dtype in (torch.float32, torch.bfloat16):
    Bm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
    Cm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
    Dv = torch.ones(D, device="cuda")
    u = torch.randn(R, D, L, generator=g, device="cuda").to(dtype)
    z = torch.randn(R, D, L, generator=g, device="cuda").to(dtype)
    A = -0.5 * torch.rand(D, N, device="cuda")
    delta = (0.5 * torch.rand(R, D, L, device="cuda")).to(dtype)
    bias = 0.5 * torch.rand(D, device="cuda")
    for use_z in (True, False):
        for sp in (True, False):
            med, best = timeit(lambda: selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, z if use_z else None, bias, sp,
                                                               need_out=not use_z, need_x=False), flush=flush)
            mufu = 16 + (2 if use_z else 0) + (2 if sp else 0) - 4 * int(os.environ.get("DIMSUM_SCAN_POLY", "0")) // 2
            print(f"poly={os.environ.get('DIMSUM_SCAN_POLY')} {str(dtype):15s} z={int(use_z)} softplus={int(sp)} mufu/elem={mufu:2d}  {med:.3f} ms")
