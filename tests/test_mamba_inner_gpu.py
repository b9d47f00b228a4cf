"""GPU parity of the fused composite (conv -> x_proj -> dt_proj -> scan -> out_proj) with the token-order gathers."""
import pytest
import torch

from golden_io import load, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


def test_mamba_inner_fn_matches_reference_golden_with_jpeg_order():
    from dimsum_b200 import mamba_inner_fn, mamba_inner_fn_cond, scanning_orders as so
    c = load("mamba_inner.npz")
    names = "xz conv_w conv_b x_proj_w dt_proj_w out_proj_w A D delta_bias".split()
    t = {n: c[n].cuda().requires_grad_(True) for n in names}
    perm, rev = c["perm"].cuda(), c["perm_rev"].cuda()
    xz_p = torch.gather(t["xz"], 2, perm[None, None, :].expand_as(t["xz"]))
    out = mamba_inner_fn(xz_p, t["conv_w"], t["conv_b"], t["x_proj_w"], t["dt_proj_w"], t["out_proj_w"], None, t["A"], None,
                         None, t["D"], delta_bias=t["delta_bias"], delta_softplus=True)
    out = torch.gather(out, 1, rev[None, :, None].expand_as(out))
    assert rel_err(out, c["out"]) <= 1e-5, rel_err(out, c["out"])
    grads = torch.autograd.grad(out, [t[n] for n in names], c["dout"].cuda())
    for n, g in zip(names, grads):
        assert rel_err(g, c["d" + n]) <= 2e-5, (n, rel_err(g, c["d" + n]))
    # the conditional variant is numerically the same function
    with torch.no_grad():
        cond = torch.randn(2, 32, 64, device="cuda")
        out_c = mamba_inner_fn_cond(xz_p, t["conv_w"], t["conv_b"], t["x_proj_w"], t["dt_proj_w"], t["out_proj_w"], None,
                                    t["A"], None, None, t["D"], delta_bias=t["delta_bias"], delta_softplus=True,
                                    init_states=cond)
        out_c = torch.gather(out_c, 1, rev[None, :, None].expand_as(out_c))
    assert rel_err(out_c, c["out"]) <= 1e-5
