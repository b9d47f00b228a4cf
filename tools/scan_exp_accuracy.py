"""Accuracy of the experimental scan-forward variants (bf16 I/O) against the fp32-I/O kernel fed the same (bf16-rounded) values.

    DIMSUM_SCAN_EX2_F16X2=1 python tools/scan_exp_accuracy.py

Two regimes: the reference test distribution (mamba/tests/ops/test_selective_scan.py:67-95) and the Mamba init regime
(A = -(1..16), delta = softplus(dt_bias) in [1e-3, 0.1], mamba_simple.py:497-521) where slow decays a = exp(delta A) ~ 0.999
dominate.  Prints max-norm relative errors (the parity metric, tolerance 2e-2 for bf16)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from dimsum_b200 import selective_scan_cuda as ssc
    R, D, L, N = 8, 1024, 256, 16
    g = torch.Generator(device="cuda").manual_seed(0)
    for regime in ("reference-test", "mamba-init", "mamba-init-L1024"):
        if regime == "mamba-init-L1024":
            L = 1024
        u = torch.randn(R, D, L, generator=g, device="cuda").bfloat16()
        z = torch.randn(R, D, L, generator=g, device="cuda").bfloat16()
        Bm = torch.randn(R, 1, N, L, generator=g, device="cuda").bfloat16()
        Cm = torch.randn(R, 1, N, L, generator=g, device="cuda").bfloat16()
        Dv = torch.randn(D, generator=g, device="cuda")
        if regime == "reference-test":
            delta = (0.5 * torch.rand(R, D, L, generator=g, device="cuda")).bfloat16()
            A = -0.5 * torch.rand(D, N, generator=g, device="cuda")
            bias = 0.5 * torch.rand(D, generator=g, device="cuda")
        else:
            delta = (0.3 * torch.randn(R, D, L, generator=g, device="cuda")).bfloat16()
            A = -torch.arange(1, N + 1, device="cuda", dtype=torch.float32).repeat(D, 1) * (1 + 0.05 * torch.rand(D, N, generator=g, device="cuda"))
            dt = torch.exp(torch.rand(D, generator=g, device="cuda") * (torch.log(torch.tensor(0.1)) - torch.log(torch.tensor(1e-3)))
                           + torch.log(torch.tensor(1e-3)))
            bias = dt + torch.log(-torch.expm1(-dt))
        want = ssc.fwd(u.float(), delta.float(), A, Bm.float(), Cm.float(), Dv, z.float(), bias, True, need_out=False, need_x=False)[2]
        got = ssc.fwd(u, delta, A, Bm, Cm, Dv, z, bias, True, need_out=False, need_x=False)[2].float()
        err = ((got - want).abs().max() / want.abs().max()).item()
        l2 = ((got - want).norm() / want.norm()).item()
        print(f"{regime:18s} L={L:5d} bf16-I/O kernel vs fp32-I/O kernel on the same values: max-norm rel err {err:.3e}, l2 rel err {l2:.3e}"
              f"  [DIMSUM_SCAN_EX2_F16X2={os.environ.get('DIMSUM_SCAN_EX2_F16X2', '0')}]")


if __name__ == "__main__":
    main()
