"""One inference scan launch at a given shape (used under `ncu` by bench.py to MEASURE the kernel's DRAM traffic in the run).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:scan_fwd -c 1 --csv python tools/one_scan.py 512 1024 256 fp32
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    R, D, L = (int(v) for v in sys.argv[1:4])
    dtype = torch.float32 if sys.argv[4] == "fp32" else torch.bfloat16
    from dimsum_b200 import selective_scan_cuda
    N = 16
    g = torch.Generator(device="cuda").manual_seed(5)
    xz = torch.randn(2 * D, R, L, generator=g, device="cuda").to(dtype).transpose(0, 1)       # channel-major GEMM view, like the model
    delta = (0.5 * torch.rand(D, R, L, generator=g, device="cuda")).to(dtype).transpose(0, 1)
    u = torch.randn(R, D, L, generator=g, device="cuda").to(dtype)
    A = -0.5 * torch.rand(D, N, generator=g, device="cuda") - 0.05
    Bm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
    Cm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
    Dv, bias = torch.ones(D, device="cuda"), torch.rand(D, device="cuda") - 3.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush.zero_()                                           # inputs out of L2: the launch reads them from DRAM like in the step
    torch.cuda.synchronize()
    selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, xz[:, D:], bias, True, need_out=False, need_x=False)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
