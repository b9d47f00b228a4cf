"""Full-width parity against the pinned oracle: `selective_scan_fn` and `causal_conv1d_fn` at
BASELINE configs[1]'s real width -- B=8, D=2048, N=16, L=256 (256px) and L=1024 (512px), fp32 and bf16 -- forward and
EVERY gradient, with the strided layouts of the model call site (u / z halves of one xz buffer, delta with strides
(L, B*L, 1)).  Distributions follow mamba/tests/ops/test_selective_scan.py:67-95; the tolerances are the
north star's (1e-5 fp32, 2e-2 bf16), a notch tighter than the reference's own CUDA-vs-ref tolerances (:54-57).

The oracle runs on the CPU (root of trust); the same-device variant of SURVEY.md section 8c is exercised by
tests/test_model_fullsize_gpu.py.  Also here: the in-kernel `perm` route at the model's width (VERDICT: only tested at D=96).
"""
import pytest
import torch

from golden_io import rel_err, tol

pytestmark = pytest.mark.gpu

R, D, N = 8, 2048, 16


def _inputs(L, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    xz = torch.randn(R, 2 * D, L, generator=g).to(dtype)
    delta = (0.5 * torch.rand(D, R, L, generator=g)).to(dtype).transpose(0, 1)          # (R, D, L) with strides (L, R*L, 1)
    A = -0.5 * torch.rand(D, N, generator=g) - 1e-3
    Bm = torch.randn(R, 1, N, L, generator=g).to(dtype)
    Cm = torch.randn(R, 1, N, L, generator=g).to(dtype)
    Dv = torch.randn(D, generator=g)
    bias = 0.5 * torch.rand(D, generator=g)
    dout = torch.randn(R, D, L, generator=g).to(dtype)
    return xz, delta, A, Bm, Cm, Dv, bias, dout


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("L", [256, 1024])
def test_selective_scan_forward_and_all_grads_at_config2_width(dtype, L):
    from dimsum_b200 import selective_scan_fn
    from oracle import ref_ops
    xz, delta, A, Bm, Cm, Dv, bias, dout = _inputs(L, dtype, 100 + L)
    names = ["xz", "delta", "A", "B", "C", "D", "delta_bias"]

    def run(fn, dev):
        leaves = [t.to(dev).detach().clone().requires_grad_(True) for t in (xz, delta, A, Bm, Cm, Dv, bias)]
        # .clone() of the strided delta keeps its (L, R*L, 1) strides (preserve_format); assert so the test cannot
        # silently degrade to the contiguous case
        lxz, ld = leaves[0], leaves[1]
        assert ld.stride() == (L, R * L, 1)
        out, last = fn(lxz[:, :D], ld, leaves[2], leaves[3], leaves[4], leaves[5], z=lxz[:, D:], delta_bias=leaves[6],
                       delta_softplus=True, return_last_state=True)
        grads = torch.autograd.grad(out, leaves, dout.to(dev))
        return out, last, grads

    # L=256: CPU oracle (root of trust).  L=1024: the same oracle fed CUDA tensors (SURVEY.md 8c "same-device oracle") -- autograd
    # through 1024 sequential CPU steps of (8, 2048, 16) tensors takes minutes
    want, want_last, want_g = run(ref_ops.selective_scan_oracle, "cpu" if L <= 256 else "cuda")
    got, got_last, got_g = run(selective_scan_fn, "cuda")
    t = tol(dtype)
    assert got.dtype == dtype and rel_err(got, want) <= t, rel_err(got, want)
    assert rel_err(got_last, want_last) <= (1e-5 if dtype == torch.float32 else 2e-2)
    for n, a, b in zip(names, got_g, want_g):
        assert a.dtype == b.dtype and a.shape == b.shape, n
        assert rel_err(a, b) <= t, (n, rel_err(a, b))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("L", [256, 1024])
def test_causal_conv1d_forward_and_grads_at_config2_width(dtype, L):
    from dimsum_b200 import causal_conv1d_fn
    from oracle import ref_ops
    g = torch.Generator().manual_seed(200 + L)
    xz = torch.randn(R, 2 * D, L, generator=g).to(dtype)
    w, b = torch.randn(D, 4, generator=g), torch.randn(D, generator=g)
    dout = torch.randn(R, D, L, generator=g).to(dtype)

    def run(fn, dev):
        lx, lw, lb = (t.to(dev).detach().clone().requires_grad_(True) for t in (xz, w, b))
        out = fn(lx[:, :D], lw, lb, "silu")                                   # x = first half of xz (batch stride 2*D*L)
        return out, torch.autograd.grad(out, [lx, lw, lb], dout.to(dev))

    want, want_g = run(ref_ops.causal_conv1d_oracle, "cpu")
    got, got_g = run(causal_conv1d_fn, "cuda")
    t = tol(dtype)
    assert rel_err(got, want) <= t
    assert rel_err(got_g[0][:, :D], want_g[0][:, :D]) <= t
    assert torch.count_nonzero(got_g[0][:, D:]) == 0
    wt = 1e-5 if dtype == torch.float32 else 2e-2
    assert rel_err(got_g[1], want_g[1]) <= wt, rel_err(got_g[1], want_g[1])
    assert rel_err(got_g[2], want_g[2]) <= wt


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("kind", ["jpeg", "zigma", "sweep"])
def test_in_kernel_perm_route_at_model_width(dtype, kind):
    """`perm` of dimsum_causal_conv1d_fwd / dimsum_selective_scan_fwd at D=1024, L=256 (the DiM-L/2 call shape): conv reads x
    through the table, the scan reads z and writes out_z through it == gather, plain ops, gather back (mamba_simple.py:634,657)."""
    from dimsum_b200 import causal_conv1d_cuda, selective_scan_cuda, scanning_orders as so
    from oracle import ref_ops
    g = torch.Generator().manual_seed(7)
    Rr, Dd, L = 4, 1024, 256
    xz = torch.randn(Rr, 2 * Dd, L, generator=g).to(dtype)
    w, b = torch.randn(Dd, 4, generator=g), torch.randn(Dd, generator=g)
    delta = (0.5 * torch.rand(Rr, Dd, L, generator=g)).to(dtype)
    A = -0.5 * torch.rand(Dd, N, generator=g) - 1e-3
    Bm, Cm = torch.randn(Rr, 1, N, L, generator=g).to(dtype), torch.randn(Rr, 1, N, L, generator=g).to(dtype)
    Dv, bias = torch.randn(Dd, generator=g), 0.5 * torch.rand(Dd, generator=g)
    table = {"jpeg": so.jpeg_zigzag, "zigma": so.zigma_path, "sweep": so.sweep_path}[kind](16)[3]
    perm = torch.from_numpy(table).long()
    rev = torch.from_numpy(so.reverse_permut_np(table)).long()
    xz_p = xz[:, :, perm]
    u_want = ref_ops.causal_conv1d_oracle(xz_p[:, :Dd], w, b, "silu")
    y_want = ref_ops.selective_scan_oracle(u_want, delta, A, Bm, Cm, Dv, z=xz_p[:, Dd:], delta_bias=bias, delta_softplus=True)
    y_want = y_want[:, :, rev]                                                # back to natural token order
    xz_d, p32 = xz.cuda(), perm.to(torch.int32).cuda()
    u_got = causal_conv1d_cuda.causal_conv1d_fwd(xz_d[:, :Dd], w.cuda(), b.cuda(), True, perm=p32)
    assert rel_err(u_got, u_want) <= tol(dtype)
    _, _, y_got = selective_scan_cuda.fwd(u_want.cuda(), delta.cuda(), A.cuda(), Bm.cuda(), Cm.cuda(), Dv.cuda(), xz_d[:, Dd:],
                                          bias.cuda(), True, need_out=False, need_x=False, perm=p32)
    assert rel_err(y_got, y_want) <= tol(dtype), rel_err(y_got, y_want)
