"""GPU parity of the inference glue kernels (modulate / gated residual with folded token order, add + RMSNorm)."""
import math

import pytest
import torch

from golden_io import rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_modulate_and_gate_residual_with_order(dtype):
    from dimsum_b200 import fused, scanning_orders as so
    g = torch.Generator(device="cuda").manual_seed(0)
    R, L, C = 5, 256, 512
    hidden = torch.randn(R, L, 2 * C, generator=g, device="cuda").to(dtype)
    x = hidden[:, :, C:]                                                   # strided half, like x2 = hidden.chunk(2, dim=2)[1]
    ada = torch.randn(R, 3 * C, generator=g, device="cuda").to(dtype)
    shift, scale, gate = ada.chunk(3, dim=1)
    order = torch.from_numpy(so.implicit_order(16, True, True)).cuda()
    inv = torch.from_numpy(so.reverse_permut_np(order.cpu().numpy())).cuda()
    want = (x.float() * (1 + scale.float().unsqueeze(1)) + shift.float().unsqueeze(1))[:, order]
    got = fused.modulate(x, shift, scale, order.to(torch.int32))
    assert rel_err(got, want) <= (1e-6 if dtype == torch.float32 else 1e-2)
    m = torch.randn(R, L, C, generator=g, device="cuda").to(dtype)
    want = x.float() + gate.float().unsqueeze(1) * m.float()[:, inv]
    got = fused.gate_residual(x, gate, m, inv.to(torch.int32))
    assert rel_err(got, want) <= (1e-6 if dtype == torch.float32 else 1e-2)
    assert torch.equal(fused.modulate(x, shift, scale), fused.modulate(x.contiguous(), shift, scale))
    # mixed precision as it occurs under autocast: fp32 residual stream, bf16 adaLN / branch outputs, bf16 GEMM input
    xf, ab = hidden.float()[:, :, C:], ada.bfloat16()
    sh, sc, gt = ab.chunk(3, dim=1)
    got = fused.modulate(xf, sh, sc, order.to(torch.int32), out_dtype=torch.bfloat16)
    want = (xf * (1 + sc.float().unsqueeze(1)) + sh.float().unsqueeze(1))[:, order]
    assert got.dtype == torch.bfloat16 and rel_err(got, want) <= 1e-2
    got = fused.gate_residual(xf, gt, m.bfloat16(), inv.to(torch.int32))
    want = xf + gt.float().unsqueeze(1) * m.bfloat16().float()[:, inv]
    assert got.dtype == torch.float32 and rel_err(got, want) <= 1e-6


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [64, 1024])
def test_add_rmsnorm_matches_oracle(dtype, C):
    from dimsum_b200 import fused
    from oracle import ref_ops
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 50, C, generator=g).to(dtype)
    res = torch.randn(3, 50, C, generator=g)
    w = 1 + 0.1 * torch.randn(C, generator=g)
    want_y, want_res = ref_ops.rms_norm_oracle(x, w, residual=res, eps=1e-5, prenorm=True)
    y, r = fused.add_rmsnorm(x.cuda(), res.cuda(), w.cuda(), 1e-5)
    assert rel_err(y, want_y) <= (2e-6 if dtype == torch.float32 else 1e-2)
    assert rel_err(r, want_res) <= 1e-6
    y2, none = fused.add_rmsnorm(x.cuda(), None, w.cuda(), 1e-5, want_residual=False)
    assert none is None and rel_err(y2, ref_ops.rms_norm_oracle(x, w, eps=1e-5)) <= (2e-6 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gelu_mul(dtype):
    from dimsum_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(2)
    x = (3 * torch.randn(7, 33, 2 * 256, generator=g, device="cuda")).to(dtype)
    a, b = x.float().chunk(2, dim=-1)
    want = torch.nn.functional.gelu(a, approximate="tanh") * b
    assert rel_err(fused.gelu_mul(x), want) <= (2e-6 if dtype == torch.float32 else 1e-2)


@pytest.mark.parametrize("layer_norm", [False, True])
@pytest.mark.parametrize("x_dtype,aux_dtype,out_dtype", [(torch.float32, torch.float32, None),
                                                         (torch.bfloat16, torch.bfloat16, torch.bfloat16),
                                                         (torch.float32, torch.bfloat16, torch.bfloat16)])
def test_norm_modulate(layer_norm, x_dtype, aux_dtype, out_dtype):
    """hidden + x -> RMSNorm / LayerNorm -> adaLN modulate in one pass vs the three reference steps
    (models_dim.py:1509-1512, :1079-1098), including the mixed precisions autocast produces."""
    import torch.nn.functional as F
    from dimsum_b200 import fused
    from oracle import ref_ops
    g = torch.Generator().manual_seed(3)
    B, L, C = 3, 37, 1024
    x = torch.randn(B, L, C, generator=g).to(x_dtype)
    res = torch.randn(B, L, C, generator=g)
    w = 1 + 0.1 * torch.randn(C, generator=g)
    ada = (0.5 * torch.randn(B, 2 * C, generator=g)).to(aux_dtype)
    shift, scale = ada.chunk(2, dim=1)                                         # row-strided views, as in the model
    for residual in (res, None):
        h = x.float() + (residual if residual is not None else 0)
        n = F.layer_norm(h, (C,), eps=1e-6) if layer_norm else ref_ops.rms_norm_oracle(h, w, eps=1e-6)
        want = n * (1 + scale.float().unsqueeze(1)) + shift.float().unsqueeze(1)
        y, r = fused.norm_modulate(x.cuda(), residual.cuda() if residual is not None else None, None if layer_norm else w.cuda(),
                                   1e-6, shift.cuda(), scale.cuda(), layer_norm=layer_norm, out_dtype=out_dtype,
                                   want_residual=residual is not None)
        assert y.dtype == (out_dtype or x_dtype)
        assert rel_err(y, want) <= (3e-6 if y.dtype == torch.float32 else 1e-2)
        if residual is not None:
            assert r.dtype == torch.float32 and rel_err(r, h) <= 1e-6
        else:
            assert r is None


def test_norm_modulate_rejects_bad_arguments():
    from dimsum_b200 import fused
    x = torch.randn(2, 8, 64, device="cuda")
    sh = torch.randn(2, 64, device="cuda")
    with pytest.raises(RuntimeError):
        fused.norm_modulate(x, x.half(), torch.ones(64, device="cuda"), 1e-5, sh, sh)          # residual must be fp32
    with pytest.raises(RuntimeError):
        fused.norm_modulate(x, None, torch.ones(64, device="cuda"), 1e-5, sh[:1], sh[:1])      # one shift row per batch row
    with pytest.raises(NotImplementedError):
        big = torch.randn(1, 2, 4096, device="cuda")
        fused.norm_modulate(big, None, torch.ones(4096, device="cuda"), 1e-5, big[:, 0], big[:, 0])   # > 1024 fp32 channels


@pytest.mark.parametrize("x_dtype,aux_dtype", [(torch.float32, torch.float32), (torch.float32, torch.bfloat16),
                                               (torch.bfloat16, torch.bfloat16)])
def test_training_glue_functions_match_autograd(x_dtype, aux_dtype):
    """modulate_fn / gate_residual_fn / gelu_mul_fn: values and every gradient against the PyTorch expressions they replace in
    the recorded pass (models_dim.py:34-35, 1510-1512; mlp.py:65-70), including the mixed dtypes autocast produces."""
    import torch.nn.functional as F
    from dimsum_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(4)
    B, L, C = 3, 50, 256
    low = x_dtype != torch.float32 or aux_dtype != torch.float32
    tol_v, tol_g = (1e-2, 2e-2) if low else (1e-6, 2e-5)

    def leaf(*shape, dtype):
        return torch.randn(*shape, generator=g, device="cuda").to(dtype).requires_grad_(True)

    x, m = leaf(B, L, C, dtype=x_dtype), leaf(B, L, C, dtype=aux_dtype)
    ada = leaf(B, 3 * C, dtype=aux_dtype)
    for fn_fused, fn_ref in ((lambda sh, sc, gt: fused.modulate_fn(x, sh, sc), lambda sh, sc, gt: x * (1 + sc.unsqueeze(1)) + sh.unsqueeze(1)),
                             (lambda sh, sc, gt: fused.gate_residual_fn(x, gt, m), lambda sh, sc, gt: x + gt.unsqueeze(1) * m)):
        outs = []
        for fn in (fn_fused, fn_ref):
            for t in (x, m, ada):
                t.grad = None
            y = fn(*ada.chunk(3, dim=1))
            gy = torch.randn(y.shape, generator=torch.Generator(device="cuda").manual_seed(9), device="cuda").to(y.dtype)
            y.backward(gy)
            outs.append((y.detach(), x.grad.clone(), ada.grad.clone(), None if m.grad is None else m.grad.clone()))
        (y1, gx1, ga1, gm1), (y2, gx2, ga2, gm2) = outs
        assert y1.dtype == y2.dtype and rel_err(y1, y2) <= tol_v
        assert gx1.dtype == gx2.dtype and rel_err(gx1, gx2) <= tol_g
        assert ga1.dtype == ga2.dtype and rel_err(ga1, ga2) <= tol_g
        if gm2 is not None:
            assert gm1.dtype == gm2.dtype and rel_err(gm1, gm2) <= tol_g
    x12 = leaf(B * L, 2 * C, dtype=x_dtype)
    outs = []
    for fn in (fused.gelu_mul_fn, lambda t: F.gelu(t[:, :C], approximate="tanh") * t[:, C:]):
        x12.grad = None
        y = fn(x12)
        y.backward(torch.ones_like(y) * 0.5)
        outs.append((y.detach(), x12.grad.clone()))
    assert rel_err(outs[0][0], outs[1][0]) <= tol_v and rel_err(outs[0][1], outs[1][1]) <= tol_g


@pytest.mark.parametrize("x_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,with_res", [(1024, True), (512, False), (64, True)])
def test_add_rmsnorm_fn_gradients_match_autograd(x_dtype, C, with_res):
    """add_rmsnorm_fn (prenorm residual stream in fp32) against autograd through the rms_norm_ref maths (layernorm.py:32-47)."""
    from dimsum_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(5)
    B, L = 3, 41
    x = torch.randn(B, L, C, generator=g, device="cuda").to(x_dtype).requires_grad_(True)
    res = torch.randn(B, L, C, generator=g, device="cuda").requires_grad_(True) if with_res else None
    w = (1 + 0.1 * torch.randn(C, generator=g, device="cuda")).requires_grad_(True)
    gy = torch.randn(B, L, C, generator=g, device="cuda").to(x_dtype)
    gh = torch.randn(B, L, C, generator=g, device="cuda")

    def ref():
        h = x.float() + (res if res is not None else 0)
        y = (h * torch.rsqrt(h.square().mean(-1, keepdim=True) + 1e-5) * w).to(x_dtype)
        return y, h

    outs = []
    for fn in (lambda: fused.add_rmsnorm_fn(x, res, w, 1e-5), ref):
        for t in (x, res, w):
            if t is not None:
                t.grad = None
        y, h = fn()
        torch.autograd.backward([y, h], [gy, gh])
        outs.append((y.detach(), h.detach(), x.grad.clone(), None if res is None else res.grad.clone(), w.grad.clone()))
    low = x_dtype != torch.float32
    for name, a, b in zip(("y", "h", "dx", "dres", "dw"), *outs):
        if b is None:
            assert a is None
            continue
        assert a.dtype == b.dtype and a.shape == b.shape, name
        assert rel_err(a, b) <= (2e-2 if low and name in ("y", "dx") else 1e-2 if low else 3e-5), (name, rel_err(a, b))


@pytest.mark.parametrize("x_dtype", [torch.float32, torch.bfloat16])
def test_gate_residual_fn_small_bf16_gates_keep_their_gradient(x_dtype):
    """adaLN gates under bf16 autocast are bf16 and small (zero-initialised, |gate| ~ 1e-3 .. 5e-2 early in training).  The
    backward computes gy * gate as gy * (1 + (gate - 1)); `gate - 1` must not be rounded to bf16 (spacing 2^-8 near -1), or
    the gradient into the mixer / MLP branch is quantised to multiples of 0.0039 -- zero for |gate| < 2e-3.  Checked per
    element, relative to the exact gy * gate, over gates spanning [1e-4, 5e-2] of both signs."""
    from dimsum_b200 import fused
    B, L, C = 2, 16, 256
    g = torch.Generator(device="cuda").manual_seed(21)
    mags = torch.logspace(-4, math.log10(5e-2), C, device="cuda")
    signs = torch.where(torch.arange(C, device="cuda") % 2 == 0, 1.0, -1.0)
    gate = (mags * signs).expand(B, C).to(torch.bfloat16).contiguous().requires_grad_(True)
    x = torch.randn(B, L, C, generator=g, device="cuda").to(x_dtype).requires_grad_(True)
    m = torch.randn(B, L, C, generator=g, device="cuda").to(torch.bfloat16).requires_grad_(True)
    y = fused.gate_residual_fn(x, gate, m)
    gy = (1.0 + torch.rand(y.shape, generator=g, device="cuda")).to(y.dtype)          # magnitudes in [1, 2): no tiny denominators
    y.backward(gy)
    want = gy.float() * gate.detach().float().unsqueeze(1)                            # exact product of the stored values
    rel = ((m.grad.float() - want).abs() / want.abs()).max().item()
    assert rel <= 2 ** -7, rel                                                       # one bf16 rounding of the result (2^-8) + slack
    assert torch.count_nonzero(m.grad) == m.grad.numel()
    want_gate = (gy.float() * m.detach().float()).sum(1)
    assert rel_err(gate.grad, want_gate) <= 1e-2
    assert rel_err(x.grad, gy) <= 1e-6


@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_cfg_euler_step_matches_forward_with_cfg_plus_euler(out_dtype):
    """dimsum_cfg_euler_step == the guidance combine of DiM.forward_with_cfg (models_dim.py:1886-1902) followed by the fixed-grid
    Euler update (integrators.py:98-111), bit for bit in fp32 (same order of roundings)."""
    from dimsum_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(8)
    n, C, H = 6, 4, 32
    half = torch.randn(n, C, H, H, generator=g, device="cuda")
    x = torch.cat([half, half])
    out = torch.randn(2 * n, C, H, H, generator=g, device="cuda").to(out_dtype)
    dt = torch.tensor(1.0 / 249, device="cuda")
    v_buf = torch.empty_like(x)
    got = fused.cfg_euler_step(x, out, 4.0, dt, v_out=v_buf)
    o = out.float()
    cond, uncond = o[:n], o[n:]
    v = uncond + 4.0 * (cond - uncond)
    want = x + dt * torch.cat([v, v])
    assert torch.equal(got, want)
    assert torch.equal(got[:n], got[n:])
    assert torch.equal(v_buf, torch.cat([v, v]))
    with pytest.raises(RuntimeError):
        fused.cfg_euler_step(x.half(), out, 4.0, dt)
