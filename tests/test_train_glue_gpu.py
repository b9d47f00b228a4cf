"""Training-path glue on the GPU: the order-folding backward of modulate / gated residual, the row table of the column-sum
kernel, the fp32-out residual RMSNorm, and the bf16 weight shadows of dimsum_b200/amp.py -- each against the PyTorch
expressions it replaces in the recorded pass (reference: dimsum/models_dim.py:34-35, 1498-1524, 1509-1512; train.py:302-321
under autocast)."""
import pytest
import torch
import torch.nn.functional as F

from golden_io import rel_err

pytestmark = pytest.mark.gpu


def _perm(L, seed):
    g = torch.Generator().manual_seed(seed)
    order = torch.randperm(L, generator=g).to(torch.int32).cuda()
    inv = torch.empty_like(order)
    inv[order.long()] = torch.arange(L, dtype=torch.int32, device="cuda")
    return order, inv


@pytest.mark.parametrize("x_dtype,aux_dtype,out_dtype", [(torch.float32, torch.float32, None),
                                                         (torch.float32, torch.bfloat16, torch.bfloat16),
                                                         (torch.float32, torch.bfloat16, None)])
def test_glue_functions_with_a_token_order_match_explicit_gathers(x_dtype, aux_dtype, out_dtype):
    """modulate_fn(x, ..., idx, inv) == modulate(x)[:, idx] and gate_residual_fn(x, gate, m, idx, inv) == x + gate * m[:, idx],
    values and every gradient; x is the strided half of a wider tensor, as in DiMBlockCombined."""
    from dimsum_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(11)
    B, L, C = 3, 64, 128
    order, inv = _perm(L, 3)
    low = aux_dtype != torch.float32
    tol_v, tol_g = (1e-2, 2e-2) if low else (1e-6, 2e-5)
    wide = torch.randn(B, L, 2 * C, generator=g, device="cuda").to(x_dtype).requires_grad_(True)
    m = torch.randn(B, L, C, generator=g, device="cuda").to(aux_dtype).requires_grad_(True)
    ada = torch.randn(B, 3 * C, generator=g, device="cuda").to(aux_dtype).requires_grad_(True)
    o, i = order.long(), inv.long()
    cases = (
        (lambda x, sh, sc, gt: fused.modulate_fn(x, sh, sc, order, inv, out_dtype=out_dtype),
         lambda x, sh, sc, gt: (x * (1 + sc.unsqueeze(1)) + sh.unsqueeze(1))[:, o]),
        (lambda x, sh, sc, gt: fused.gate_residual_fn(x, gt, m, inv, order, out_dtype=out_dtype),
         lambda x, sh, sc, gt: x + gt.unsqueeze(1) * m[:, i]),
    )
    for fn_fused, fn_ref in cases:
        outs = []
        for fn in (fn_fused, fn_ref):
            for t in (wide, m, ada):
                t.grad = None
            y = fn(wide[:, :, C:], *ada.chunk(3, dim=1))
            gy = torch.randn(y.shape, generator=torch.Generator(device="cuda").manual_seed(9), device="cuda")
            (y.float() * gy).sum().backward()
            outs.append((y.detach(), wide.grad.clone(), ada.grad.clone(), None if m.grad is None else m.grad.clone()))
        (y1, gx1, ga1, gm1), (y2, gx2, ga2, gm2) = outs
        assert y1.dtype == (out_dtype or y2.dtype)
        assert rel_err(y1, y2) <= tol_v, rel_err(y1, y2)
        assert gx1.dtype == gx2.dtype and rel_err(gx1, gx2) <= tol_g, rel_err(gx1, gx2)
        assert ga1.dtype == ga2.dtype and rel_err(ga1, ga2) <= tol_g, rel_err(ga1, ga2)
        if gm2 is not None:
            assert gm1.dtype == gm2.dtype and rel_err(gm1, gm2) <= tol_g, rel_err(gm1, gm2)
    with pytest.raises(RuntimeError):
        fused.modulate_fn(wide[:, :, C:], ada[:, :C], ada[:, C:2 * C], order, None)


def test_token_colsum_pairs_rows_through_the_table():
    from dimsum_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(2)
    B, L, C = 4, 100, 64
    order, _ = _perm(L, 5)
    gr = torch.randn(B, L, C, generator=g, device="cuda").to(torch.bfloat16)
    x = torch.randn(B, L, C, generator=g, device="cuda")
    sg, sgx = fused.token_colsum(gr, x, out_dtype=torch.float32, x_idx=order)
    assert rel_err(sg, gr.float().sum(1)) <= 1e-6
    assert rel_err(sgx, (gr.float() * x[:, order.long()]).sum(1)) <= 1e-5
    _, plain = fused.token_colsum(gr, x, want_sum_g=False, out_dtype=torch.float32)
    assert rel_err(plain, (gr.float() * x).sum(1)) <= 1e-5


def test_add_rmsnorm_fn_emits_fp32_rows_from_a_bf16_input():
    """hidden = hidden + x; norm_2(hidden) with x the bf16 output of a GEMM under autocast (models_dim.py:1509-1511): the sum
    and the normalised rows stay fp32, the gradient of x comes back in bf16."""
    from dimsum_b200 import fused
    g = torch.Generator(device="cuda").manual_seed(8)
    rows, C = 70, 1024
    x = torch.randn(2, rows, C, generator=g, device="cuda").to(torch.bfloat16).requires_grad_(True)
    res = torch.randn(2, rows, C, generator=g, device="cuda").requires_grad_(True)
    w = (1 + 0.1 * torch.randn(C, generator=g, device="cuda")).requires_grad_(True)
    gy = torch.randn(2, rows, C, generator=g, device="cuda")
    gh = torch.randn(2, rows, C, generator=g, device="cuda")
    outs = []
    for fused_path in (True, False):
        for t in (x, res, w):
            t.grad = None
        if fused_path:
            y, h = fused.add_rmsnorm_fn(x, res, w, 1e-5, out_dtype=torch.float32)
        else:
            h = x.float() + res
            y = h * torch.rsqrt(h.square().mean(-1, keepdim=True) + 1e-5) * w
        ((y * gy).sum() + (h * gh).sum()).backward()
        outs.append((y.detach(), h.detach(), x.grad.clone(), res.grad.clone(), w.grad.clone()))
    a, b = outs
    assert a[0].dtype == torch.float32 and a[1].dtype == torch.float32 and a[2].dtype == torch.bfloat16
    assert rel_err(a[0], b[0]) <= 1e-5 and rel_err(a[1], b[1]) <= 1e-6
    assert rel_err(a[2], b[2]) <= 1e-2 and rel_err(a[3], b[3]) <= 2e-5 and rel_err(a[4], b[4]) <= 2e-5


def test_weight_shadows_reproduce_autocast():
    """Linear / transposed in_proj on the bf16 shadows: the forward and the input gradient are the bits autocast produces, the
    weight and bias gradients arrive in fp32 and agree with autocast's (which are rounded to bf16 on the way) to bf16 accuracy;
    a shadow older than its master is refreshed on use."""
    from dimsum_b200 import amp
    from dimsum_b200.models_dim import Linear
    torch.manual_seed(0)
    lin = Linear(256, 384, bias=True).cuda()
    x = torch.randn(4, 96, 256, device="cuda", requires_grad=True)
    gy = torch.randn(4, 96, 384, device="cuda")

    def run():
        for t in (x, lin.weight, lin.bias):
            t.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = lin(x)
            z = amp.weight_times_rows_t(lin.weight, x.reshape(-1, 256))
        ((y.float() * gy).sum() + z.float().square().sum() * 1e-3).backward()
        return y.detach(), z.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()

    plain = run()
    sh = amp.Bf16Shadows(lin)
    shadowed = run()
    assert torch.equal(plain[0], shadowed[0]) and torch.equal(plain[1], shadowed[1])
    assert rel_err(shadowed[2], plain[2]) <= 1e-2
    assert shadowed[3].dtype == torch.float32 and rel_err(shadowed[3], plain[3]) <= 1e-2
    assert shadowed[4].dtype == torch.float32 and rel_err(shadowed[4], plain[4]) <= 1e-2
    # fp32 reference of the same bf16-rounded operands: the shadow path's weight gradient skips one rounding
    w16, x16 = lin.weight.detach().bfloat16().float(), x.detach().bfloat16().float()
    gyy = gy.bfloat16().float().reshape(-1, 384)
    gz = (2e-3 * (w16 @ x16.reshape(-1, 256).t()).bfloat16().float()).bfloat16().float()
    want_gw = gyy.t() @ x16.reshape(-1, 256) + gz @ x16.reshape(-1, 256)
    assert rel_err(shadowed[3], want_gw) <= rel_err(plain[3], want_gw) + 1e-6
    with torch.no_grad():                      # optimizer-style in-place update without refresh(): picked up on the next call
        lin.weight.mul_(0.5)
    after = run()
    assert rel_err(after[0].float(), 0.5 * (plain[0].float() - lin.bias.detach().bfloat16().float())
                   + lin.bias.detach().bfloat16().float()) <= 2e-2
    sh.refresh()
    assert torch.equal(lin.weight._dimsum_bf16, lin.weight.detach().bfloat16())
    sh.detach()
    assert not hasattr(lin.weight, "_dimsum_bf16")
    with torch.autocast("cuda", dtype=torch.bfloat16):            # no shadows, no autocast: plain F.linear
        assert torch.equal(lin(x), F.linear(x, lin.weight, lin.bias))


def test_train_step_with_shadows_and_fused_clip_tracks_plain_autocast():
    """Three optimizer steps of a small DiM under bf16 autocast: plain autocast + clip_grad_norm_, against the weight shadows
    and against the clip folded into the fused AdamW kernel -- same losses and weights to bf16 accuracy, and gradients that
    end up clipped either way (tools/train_step.py, eager)."""
    import os, sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    from train_step import TrainStep
    dev = torch.device("cuda", 0)
    runs = []
    for shadows, fuse_clip in ((False, False), (True, False), (True, True)):
        ts = TrainStep(dev, 0, 1, batch=4, dtype="bf16", depth=4, use_graph=False, shadows=shadows, fuse_clip=fuse_clip)
        assert (ts.shadows is not None) == shadows
        batch = [b.clone() for b in ts.draw()]
        losses = [float(ts.step(batch).detach()) for _ in range(3)]
        w = torch.cat([p.detach().flatten()[:4096] for n, p in ts.model.named_parameters() if "cond_proj" not in n])
        gnorm = torch.nn.utils.get_total_norm([p.grad for p in ts.model.parameters() if p.grad is not None])
        runs.append((losses, w, ts.missing_grads(), float(gnorm)))
    (l0, w0, m0, g0) = runs[0]
    for l1, w1, m1, g1 in runs[1:]:
        assert m0 == [] and m1 == []
        assert all(abs(a - b) <= 2e-2 * abs(a) for a, b in zip(l0, l1)), (l0, l1)
        assert rel_err(w1, w0) <= 1e-3, rel_err(w1, w0)
        assert g1 <= 1.0 + 1e-3 and abs(g1 - g0) <= 2e-2 * g0, (g0, g1)
