"""GPU: edge cases of the hot path against the pinned oracle -- empty batch, single element, ragged / unaligned lengths,
channel counts that do not fill a CTA slab, small state counts, 16-bit I/O on unaligned rows."""
import pytest
import torch

from golden_io import rel_err, tol

pytestmark = pytest.mark.gpu


def _scan_inputs(R, D, L, N, dtype, seed, groups=0):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(R, D, L, generator=g).to(dtype)
    delta = (0.5 * torch.rand(R, D, L, generator=g)).to(dtype)
    A = -0.5 * torch.rand(D, N, generator=g) - 0.01
    shape = (R, N, L) if groups == 0 else (R, groups, N, L)
    Bm, Cm = torch.randn(shape, generator=g).to(dtype), torch.randn(shape, generator=g).to(dtype)
    Dv, z, bias = torch.randn(D, generator=g), torch.randn(R, D, L, generator=g).to(dtype), torch.rand(D, generator=g) - 1.0
    return u, delta, A, Bm, Cm, Dv, z, bias


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(1, 1, 1, 1), (3, 130, 5, 3), (2, 64, 16, 16), (2, 257, 33, 16), (1, 40, 47, 7), (2, 12, 31, 8)])
def test_scan_forward_and_backward_odd_shapes(dtype, shape):
    from dimsum_b200 import selective_scan_fn
    from oracle import ref_ops
    R, D, L, N = shape
    groups = 2 if D % 2 == 0 and N == 8 else 0
    args = _scan_inputs(R, D, L, N, dtype, seed=sum(shape), groups=groups)
    cpu = [t.clone().requires_grad_(True) for t in args]
    want, last_w = ref_ops.selective_scan_oracle(*cpu[:6], z=cpu[6], delta_bias=cpu[7], delta_softplus=True, return_last_state=True)
    gout = torch.randn(want.shape, generator=torch.Generator().manual_seed(1)).to(dtype)
    gw = torch.autograd.grad(want, cpu, gout)
    dev = [t.detach().cuda().requires_grad_(True) for t in args]
    got, last = selective_scan_fn(*dev[:6], z=dev[6], delta_bias=dev[7], delta_softplus=True, return_last_state=True)
    assert rel_err(got, want) <= tol(dtype)
    assert rel_err(last, last_w) <= 1e-5
    gg = torch.autograd.grad(got, dev, gout.cuda())
    for name, a, b in zip("u delta A B C D z delta_bias".split(), gg, gw):
        assert a.shape == b.shape
        if b.abs().max() > 0:
            assert rel_err(a, b) <= (2e-5 if dtype == torch.float32 else 3e-2), (name, rel_err(a, b))


def test_empty_batch_is_a_no_op():
    from dimsum_b200 import causal_conv1d_fn, selective_scan_fn, wavelet_packet
    u = torch.randn(0, 8, 16, device="cuda")
    A = -torch.rand(8, 4, device="cuda")
    Bm = torch.randn(0, 4, 16, device="cuda")
    assert selective_scan_fn(u, u, A, Bm, Bm).shape == (0, 8, 16)
    assert causal_conv1d_fn(u, torch.randn(8, 4, device="cuda")).shape == (0, 8, 16)
    assert wavelet_packet(torch.randn(0, 64, 16, device="cuda")).shape == (0, 64, 16)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 3, 2), (3, 5, 7), (2, 129, 100), (1, 2, 2049)])
@pytest.mark.parametrize("width", [2, 4])
def test_conv_odd_shapes_forward_backward(dtype, shape, width):
    from dimsum_b200 import causal_conv1d_fn
    from oracle import ref_ops
    R, D, L = shape
    g = torch.Generator().manual_seed(R * 100 + D * 10 + L + width)
    x = torch.randn(R, D, L, generator=g).to(dtype)
    w, b = torch.randn(D, width, generator=g), torch.randn(D, generator=g)
    cpu = [t.clone().requires_grad_(True) for t in (x, w, b)]
    want = ref_ops.causal_conv1d_oracle(*cpu, "silu")
    gout = torch.randn(want.shape, generator=g).to(dtype)
    gw = torch.autograd.grad(want, cpu, gout)
    dev = [t.detach().cuda().requires_grad_(True) for t in (x, w, b)]
    got = causal_conv1d_fn(*dev, "silu")
    assert rel_err(got, want) <= tol(dtype)
    gg = torch.autograd.grad(got, dev, gout.cuda())
    for a, bb in zip(gg, gw):
        assert rel_err(a, bb) <= (2e-5 if dtype == torch.float32 else 3e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_scan_and_conv_through_every_order_family(dtype):
    """conv(perm) -> scan(perm) equals gather, plain ops, gather back, for a table of each family (a10)."""
    from dimsum_b200 import causal_conv1d_cuda, selective_scan_cuda, scanning_orders as so
    from oracle import ref_ops
    R, D, L, N = 2, 96, 256, 16
    u, delta, A, Bm, Cm, Dv, z, bias = _scan_inputs(R, D, L, N, dtype, seed=9)
    w, cb = torch.randn(D, 4, generator=torch.Generator().manual_seed(3)), torch.randn(D, generator=torch.Generator().manual_seed(4))
    for family, which in (("sweep", 3), ("zigma", 6), ("jpeg", 1)):
        perm = torch.from_numpy(so.SCAN_ZOO[family](16)[which])
        inv = torch.from_numpy(so.reverse_permut_np(perm.numpy()))
        xc = ref_ops.causal_conv1d_oracle(u[:, :, perm], w, cb, "silu")
        want = ref_ops.selective_scan_oracle(xc, delta, A, Bm, Cm, Dv, z=z[:, :, perm], delta_bias=bias, delta_softplus=True)[:, :, inv]
        p32 = perm.to(torch.int32).cuda()
        xd = causal_conv1d_cuda.causal_conv1d_fwd(u.cuda(), w.cuda(), cb.cuda(), True, perm=p32)
        assert rel_err(xd, xc) <= tol(dtype)
        _, _, got = selective_scan_cuda.fwd(xc.cuda(), delta.cuda(), A.cuda(), Bm.unsqueeze(1).cuda(), Cm.unsqueeze(1).cuda(), Dv.cuda(),
                                            z.cuda(), bias.cuda(), True, need_out=False, need_x=False, perm=p32)
        assert rel_err(got, want) <= tol(dtype), (family, rel_err(got, want))
