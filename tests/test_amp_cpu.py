"""Host logic of the bf16 weight shadows (dimsum_b200/amp.py) under CPU autocast: same forward as autocast, gradients to the
fp32 masters, stale shadows refreshed on use, nothing attached -> plain F.linear.  (The CUDA specifics -- fp32 GEMM output for the
weight gradient, the column-sum kernel for the bias gradient -- are covered by tests/test_train_glue_gpu.py.)"""
import torch
import torch.nn.functional as F

from dimsum_b200 import amp
from dimsum_b200.models_dim import Linear, _Cond, _ada


def _run(lin, x, gy):
    for t in (x, lin.weight, lin.bias):
        t.grad = None
    with torch.autocast("cpu", dtype=torch.bfloat16):
        y = lin(x)
        z = amp.weight_times_rows_t(lin.weight, x.reshape(-1, x.shape[-1]))
    ((y.float() * gy).sum() + z.float().square().sum() * 1e-3).backward()
    return y.detach(), z.detach(), x.grad.clone(), lin.weight.grad.clone(), lin.bias.grad.clone()


def test_shadows_follow_cpu_autocast_and_their_masters():
    torch.manual_seed(0)
    lin = Linear(16, 24, bias=True)
    x = torch.randn(3, 10, 16, requires_grad=True)
    gy = torch.randn(3, 10, 24)
    plain = _run(lin, x, gy)
    sh = amp.Bf16Shadows(lin)
    assert len(sh.masters) == 2 and lin.weight._dimsum_bf16.dtype == torch.bfloat16
    got = _run(lin, x, gy)
    assert torch.equal(plain[0], got[0]) and torch.equal(plain[1], got[1])
    for a, b in zip(plain[2:], got[2:]):
        assert a.dtype == b.dtype == torch.float32
        assert (a - b).abs().max() <= 2e-2 * a.abs().max()
    with torch.no_grad():                                   # an optimizer-style in-place step without refresh()
        lin.weight.mul_(2.0)
    assert lin.weight._dimsum_bf16_version != lin.weight._version
    after = _run(lin, x, gy)
    assert torch.equal(lin.weight._dimsum_bf16, lin.weight.detach().bfloat16())
    assert (after[1].float() - 2 * plain[1].float()).abs().max() <= 2e-2 * plain[1].float().abs().max()
    with torch.no_grad():
        lin.bias.add_(1.0)
    sh.refresh()
    assert torch.equal(lin.bias._dimsum_bf16, lin.bias.detach().bfloat16())
    assert lin.bias._dimsum_bf16_version == lin.bias._version
    assert torch.equal(lin(x), F.linear(x, lin.weight, lin.bias))          # no autocast: the masters, plain F.linear
    sh.detach()
    assert not hasattr(lin.weight, "_dimsum_bf16") and amp.shadow_of(lin.weight) is None


def test_adaln_heads_share_one_activation():
    torch.manual_seed(1)
    head = torch.nn.Sequential(torch.nn.SiLU(), Linear(8, 12))
    c = torch.randn(4, 8)
    cond = _Cond(c)
    assert torch.equal(cond.act, F.silu(c)) and cond.raw is c
    assert torch.equal(_ada(head, cond), head(c)) and torch.equal(_ada(head, c), head(c))
