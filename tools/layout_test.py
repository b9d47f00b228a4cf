"""Developer experiment: does the scan kernel's speed depend on the VALUES it is fed? (it should not)"""
import sys
import torch
sys.path.insert(0, "/root/repo")
from dimsum_b200 import selective_scan_cuda
from tools.bench_ops import timeit
R, D, L, N = 512, 1024, 256, 16
g = torch.Generator(device="cuda").manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
dtype = torch.float32
Bm = torch.randn(R, 1, N, L, generator=g, device="cuda")
Cm = torch.randn(R, 1, N, L, generator=g, device="cuda")
Dv = torch.ones(D, device="cuda")
u = torch.randn(R, D, L, generator=g, device="cuda")
z = torch.randn(R, D, L, generator=g, device="cuda")
cases = {
    "A=-0.5rand delta=0.5rand bias=0.5rand": (-0.5 * torch.rand(D, N, device="cuda"), 0.5 * torch.rand(R, D, L, device="cuda"), 0.5 * torch.rand(D, device="cuda")),
    "A=-(1..16) delta=0.5rand": (-torch.arange(1, N + 1, device="cuda").float().repeat(D, 1), 0.5 * torch.rand(R, D, L, device="cuda"), 0.5 * torch.rand(D, device="cuda")),
    "A=-(1..16) delta~N(0,1) bias=-4.6": (-torch.arange(1, N + 1, device="cuda").float().repeat(D, 1), torch.randn(R, D, L, device="cuda"), torch.full((D,), -4.6, device="cuda")),
    "A=-(1..16) delta~N(0,3) bias=-4.6": (-torch.arange(1, N + 1, device="cuda").float().repeat(D, 1), 3 * torch.randn(R, D, L, device="cuda"), torch.full((D,), -4.6, device="cuda")),
    "A=-0.5rand delta=20+rand (fast decay)": (-0.5 * torch.rand(D, N, device="cuda"), 20 + torch.rand(R, D, L, device="cuda"), torch.zeros(D, device="cuda")),
    "tiny u,B,C (1e-20)": (-0.5 * torch.rand(D, N, device="cuda"), 0.5 * torch.rand(R, D, L, device="cuda"), 0.5 * torch.rand(D, device="cuda")),
}
for name, (A, delta, bias) in cases.items():
    uu, BB, CC = (u * 1e-20, Bm * 1e-10, Cm * 1e-10) if name.startswith("tiny") else (u, Bm, Cm)
    med, best = timeit(lambda: selective_scan_cuda.fwd(uu, delta, A, BB, CC, Dv, z, bias, True, need_out=False, need_x=False), flush=flush)
    print(f"{name:45s} {med:.3f} ms")
