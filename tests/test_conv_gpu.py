"""GPU parity of causal_conv1d_fn vs reference golden vectors and the pinned oracle."""
import pytest
import torch

from golden_io import load, rel_err, tol

pytestmark = pytest.mark.gpu
GOLD = load("conv.npz")


@pytest.mark.parametrize("case", sorted(GOLD))
def test_forward_and_backward_match_reference_golden(case):
    from dimsum_b200 import causal_conv1d_fn
    c = GOLD[case]
    x = c["x"].cuda().requires_grad_(True)
    w = c["weight"].cuda().requires_grad_(True)
    b = c["bias"].cuda().requires_grad_(True) if "bias" in c else None
    out = causal_conv1d_fn(x, w, b, "silu" if c["silu"] else None)
    t = tol(x.dtype)
    assert out.dtype == x.dtype and rel_err(out, c["out"]) <= t, rel_err(out, c["out"])
    grads = torch.autograd.grad(out, [x, w] + ([b] if b is not None else []), c["dout"].cuda())
    assert rel_err(grads[0], c["dx"]) <= t
    wt = 1e-5 if w.dtype == torch.float32 and x.dtype == torch.float32 else 2e-2
    assert rel_err(grads[1], c["dweight"]) <= wt, rel_err(grads[1], c["dweight"])
    if b is not None:
        assert rel_err(grads[2], c["dbias"]) <= wt


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("L", [8, 64, 151, 256, 372, 1024, 4096])
def test_strided_view_of_xz_matches_oracle(dtype, L):
    """x is the first half of xz (batch stride 2*D*L), as in MambaInnerFn (selective_scan_interface.py:833)."""
    from dimsum_b200 import causal_conv1d_fn
    from oracle import ref_ops
    g = torch.Generator().manual_seed(L)
    R, D = 2, 160
    xz = torch.randn(R, 2 * D, L, generator=g).to(dtype)
    w, b = torch.randn(D, 4, generator=g), torch.randn(D, generator=g)
    want = ref_ops.causal_conv1d_oracle(xz[:, :D], w, b, "silu")
    got = causal_conv1d_fn(xz.cuda()[:, :D], w.cuda(), b.cuda(), "silu")
    assert rel_err(got, want) <= tol(dtype)


def test_conv_through_token_order_matches_gather_then_conv():
    from dimsum_b200 import causal_conv1d_cuda, scanning_orders as so
    from oracle import ref_ops
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 96, 256, generator=g)
    w, b = torch.randn(96, 4, generator=g), torch.randn(96, generator=g)
    perm = torch.from_numpy(so.jpeg_zigzag(16)[5])
    want = ref_ops.causal_conv1d_oracle(x[:, :, perm], w, b, "silu")
    got = causal_conv1d_cuda.causal_conv1d_fwd(x.cuda(), w.cuda(), b.cuda(), True, perm=perm.to(torch.int32).cuda())
    assert rel_err(got, want) <= 1e-5


def test_fwd_cond_writes_into_the_given_buffer():
    """Reference quirk Q1: the 'conditioning' tensor is only the output buffer (causal_conv1d.cpp:326)."""
    from dimsum_b200 import causal_conv1d_cuda
    x = torch.randn(2, 8, 32, device="cuda")
    w, b = torch.randn(8, 4, device="cuda"), torch.randn(8, device="cuda")
    plain = causal_conv1d_cuda.causal_conv1d_fwd(x, w, b, True)
    for fill in (0.0, 123.0):
        buf = torch.full_like(x, fill)
        out = causal_conv1d_cuda.causal_conv1d_fwd_cond(x, w, b, True, buf)
        assert out.data_ptr() == buf.data_ptr() and torch.equal(out, plain)


def test_determinism_of_forward_and_dx():
    """Reference test_causal_conv1d_race_condition (tests/test_causal_conv1d.py:123-180), shortened."""
    from dimsum_b200 import causal_conv1d_cuda
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(4, 256, 512, generator=g, device="cuda")
    w, b = torch.randn(256, 4, generator=g, device="cuda"), torch.randn(256, generator=g, device="cuda")
    dout = torch.randn(4, 256, 512, generator=g, device="cuda")
    out0 = causal_conv1d_cuda.causal_conv1d_fwd(x, w, b, True)
    dx0, dw0, db0 = causal_conv1d_cuda.causal_conv1d_bwd(x, w, b, dout, None, True)
    for _ in range(50):
        assert torch.equal(causal_conv1d_cuda.causal_conv1d_fwd(x, w, b, True), out0)
        dx, dw, db = causal_conv1d_cuda.causal_conv1d_bwd(x, w, b, dout, None, True)
        assert torch.equal(dx, dx0)
        assert (dw - dw0).abs().max() <= 1e-3 and (db - db0).abs().max() <= 1e-3


def test_error_behaviour():
    from dimsum_b200 import causal_conv1d_fn
    x = torch.randn(1, 4, 16, device="cuda")
    with pytest.raises(RuntimeError, match="width between 2 and 4"):
        causal_conv1d_fn(x, torch.randn(4, 5, device="cuda"))
    with pytest.raises(NotImplementedError):
        causal_conv1d_fn(x, torch.randn(4, 4, device="cuda"), None, "gelu")
    with pytest.raises(NotImplementedError):
        causal_conv1d_fn(x.transpose(1, 2).contiguous().transpose(1, 2), torch.randn(4, 4, device="cuda"))  # channel-last
