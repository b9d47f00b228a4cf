"""GPU parity of the tcgen05 kernel that fuses causal conv1d + SiLU in front of the x_proj contraction
(dimsum_conv_xproj_fwd; reference: the first half of MambaInnerFn*.forward, selective_scan_interface.py:836-866):
u must equal the stand-alone conv kernel bit for bit, x_dbl must match F.linear on the oracle's conv output within the north
star's tolerances (1e-5 fp32 with the 3xTF32 split, 2e-2 for 16-bit I/O; single-pass TF32 is checked against the
TF32 error model), in both output layouts, on strided halves of xz, ragged token tiles and the `init_states` buffer."""
import pytest
import torch
import torch.nn.functional as F

from golden_io import rel_err

pytestmark = pytest.mark.gpu


def _case(R, D, L, E, dtype, seed=0, w_dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    xz = torch.randn(R, 2 * D, L, generator=g).to(dtype)
    w = torch.randn(D, 4, generator=g).to(w_dtype)
    b = torch.randn(D, generator=g).to(w_dtype)
    xw = (torch.randn(E, D, generator=g) / D ** 0.5).to(dtype)
    return xz, w, b, xw


@pytest.mark.parametrize("shape", [(4, 1024, 256, 64), (2, 128, 200, 64), (3, 256, 1024, 48), (1, 64, 8, 8), (2, 512, 132, 128)])
@pytest.mark.parametrize("mode", ["fp32-3xtf32", "fp32-tf32", "bf16", "fp16"])
def test_conv_xproj_matches_conv_then_linear(shape, mode):
    from dimsum_b200 import causal_conv1d_cuda as ccc
    from oracle import ref_ops
    R, D, L, E = shape
    dtype = {"fp32-3xtf32": torch.float32, "fp32-tf32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}[mode]
    if dtype != torch.float32 and (D % 64 or L % 8):
        pytest.skip("16-bit I/O needs dim % 64 == 0 and seqlen % 8 == 0")
    xz, w, b, xw = _case(R, D, L, E, dtype, seed=D + L)
    x_d = xz.cuda()[:, :D]                                     # first half of xz: batch stride 2*D*L
    if mode == "fp32-3xtf32" and E > 64:
        assert not ccc.conv_xproj_supported(x_d, w.cuda(), xw.cuda(), precise=True)
        pytest.skip("hi + lo weight tiles of more than 64 rows do not fit the ring")
    assert ccc.conv_xproj_supported(x_d, w.cuda(), xw.cuda())
    u, xd = ccc.conv_xproj_fwd(x_d, w.cuda(), b.cuda(), xw.cuda(), precise=(mode == "fp32-3xtf32"))
    u_ref = ccc.causal_conv1d_fwd(x_d, w.cuda(), b.cuda(), True)
    assert torch.equal(u, u_ref)                               # same arithmetic as the stand-alone conv kernel
    u_orc = ref_ops.causal_conv1d_oracle(xz[:, :D], w, b, "silu")
    want = F.linear(u_orc.double().transpose(1, 2), xw.double()).transpose(1, 2)     # (R, E, L), exact product of the stored values
    assert xd.shape == (R, E, L) and xd.dtype == dtype
    tol = {"fp32-3xtf32": 1e-5, "fp32-tf32": 2e-3, "bf16": 2e-2, "fp16": 4e-3}[mode]
    assert rel_err(xd, want.float()) <= tol, rel_err(xd, want.float())
    # split layout: dt as the (rank, R*L) operand of the dt_proj GEMM, B / C as halves of (R, E - rank, L)
    if E >= 16:
        rank = E // 2 if (E // 2) % 8 == 0 else 8
        u2, dt, bc = ccc.conv_xproj_fwd(x_d, w.cuda(), b.cuda(), xw.cuda(), precise=(mode == "fp32-3xtf32"), split=rank)
        assert torch.equal(u2, u)
        assert torch.equal(dt.view(rank, R, L).transpose(0, 1), xd[:, :rank])
        assert torch.equal(bc, xd[:, rank:])


def test_conv_xproj_writes_u_into_the_init_states_buffer_and_rejects_bad_layouts():
    from dimsum_b200 import causal_conv1d_cuda as ccc
    xz, w, b, xw = _case(2, 128, 64, 64, torch.float32)
    x_d = xz.cuda()[:, :128]
    buf = torch.full((2, 128, 64), 7.0, device="cuda")
    u, xd = ccc.conv_xproj_fwd(x_d, w.cuda(), b.cuda(), xw.cuda(), precise=True, out=buf)
    assert u.data_ptr() == buf.data_ptr() and torch.equal(buf, ccc.causal_conv1d_fwd(x_d, w.cuda(), b.cuda(), True))
    assert not ccc.conv_xproj_supported(x_d[:, :100], w.cuda()[:100], xw.cuda()[:, :100])          # dim % 32
    assert not ccc.conv_xproj_supported(x_d[:, :, :63], w.cuda(), xw.cuda())                        # seqlen % 4
    assert not ccc.conv_xproj_supported(x_d, w.cuda(), xw.cuda()[:60])                              # n_out % 8
    with pytest.raises(RuntimeError):
        ccc.conv_xproj_fwd(x_d[:, :, :63], w.cuda(), b.cuda(), xw.cuda())


@pytest.mark.parametrize("dtype,allow_tf32", [(torch.float32, False), (torch.float32, True), (torch.bfloat16, True)])
def test_mamba_inner_fn_fused_and_two_step_paths_agree(dtype, allow_tf32, monkeypatch):
    """MambaInnerFn forward + every gradient with the fused conv+x_proj kernel vs DIMSUM_FUSED_XPROJ=0 (round-1 path), at the
    DiM-L/2 mixer shape (d_inner 1024, dt_rank 32, N 16, L 256)."""
    from dimsum_b200 import mamba_inner_fn
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = allow_tf32
    try:
        g = torch.Generator(device="cuda").manual_seed(3)
        R, Dm, L, N, rank, dm = 4, 1024, 256, 16, 32, 512
        mk = lambda *s, scale=1.0: (torch.randn(*s, generator=g, device="cuda") * scale)
        xz = mk(R, 2 * Dm, L).to(dtype).requires_grad_(True)
        conv_w, conv_b = mk(Dm, 1, 4, scale=0.5).requires_grad_(True), mk(Dm, scale=0.1).requires_grad_(True)
        xw = mk(rank + 2 * N, Dm, scale=Dm ** -0.5).requires_grad_(True)
        dtw = mk(Dm, rank, scale=rank ** -0.5).requires_grad_(True)
        ow = mk(dm, Dm, scale=Dm ** -0.5).requires_grad_(True)
        A = (-torch.rand(Dm, N, generator=g, device="cuda") - 0.05).requires_grad_(True)
        Dv = torch.ones(Dm, device="cuda", requires_grad=True)
        bias = (torch.rand(Dm, generator=g, device="cuda") - 3.0).requires_grad_(True)
        leaves = [xz, conv_w, conv_b, xw, dtw, ow, A, Dv, bias]
        dout = mk(R, L, dm).to(dtype)
        res = {}
        for flag in ("1", "0"):
            monkeypatch.setenv("DIMSUM_FUSED_XPROJ", flag)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=dtype == torch.bfloat16):
                out = mamba_inner_fn(xz, conv_w, conv_b, xw, dtw, ow, None, A, None, None, Dv, delta_bias=bias, delta_softplus=True)
            res[flag] = (out.detach(), torch.autograd.grad(out, leaves, dout))
        tol = 2e-2 if dtype == torch.bfloat16 else (5e-3 if allow_tf32 else 5e-5)      # both sides are approximations of the fp32 sums
        assert rel_err(res["1"][0], res["0"][0]) <= tol, rel_err(res["1"][0], res["0"][0])
        for name, a, b in zip("xz conv_w conv_b x_proj_w dt_proj_w out_proj_w A D delta_bias".split(), res["1"][1], res["0"][1]):
            assert a.shape == b.shape and rel_err(a, b) <= tol, (name, rel_err(a, b))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
