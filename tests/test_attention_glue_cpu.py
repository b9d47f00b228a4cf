"""CPU checks of the attention glue of the recorded pass: the q/k/v split with its single-pass backward and the head-axis
concatenation of the two cross attentions against the reference expressions (dimsum/attention_fusion.py:61-84)."""
import torch
import torch.nn.functional as F

from dimsum_b200.models_dim import CrossAttentionFusion, _split_qkv


def _plain_split(qkv, H):
    B, N, C3 = qkv.shape
    return qkv.view(B, N, 3, H, C3 // 3 // H).permute(2, 0, 3, 1, 4).unbind(0)


def test_qkv_split_backward_matches_the_view_permute_unbind_chain():
    torch.manual_seed(0)
    B, N, H, D = 2, 16, 4, 8
    base = torch.randn(B, N, 3 * H * D)
    w = torch.randn(B, H, N, D)
    grads = []
    for split in (_split_qkv, _plain_split):
        qkv = base.clone().requires_grad_(True)
        q, k, v = split(qkv, H)
        assert q.shape == (B, H, N, D)
        (g,) = torch.autograd.grad((F.scaled_dot_product_attention(q, k, v) * w).sum(), qkv)
        grads.append(g)
    assert torch.equal(grads[0], grads[1])
    grads = []
    for split in (_split_qkv, _plain_split):          # v unused: its slot of the gradient is zero-filled
        qkv = base.clone().requires_grad_(True)
        q, k, v = split(qkv, H)
        (g,) = torch.autograd.grad((q * k).sum(), qkv)
        grads.append(g)
    assert torch.equal(grads[0], grads[1])
    with torch.no_grad():
        for a, b in zip(_split_qkv(base, H), _plain_split(base, H)):
            assert torch.equal(a, b)


def test_cross_attention_fusion_matches_the_reference_expression():
    torch.manual_seed(1)
    m = CrossAttentionFusion(64, num_heads=4)
    x1 = torch.randn(2, 16, 32, requires_grad=True)
    x2 = torch.randn(2, 16, 32, requires_grad=True)

    def ref(x1, x2):
        B, N, C = x1.shape
        q1, k1, v1 = _plain_split(m.qkv1(x1), m.num_heads)
        q2, k2, v2 = _plain_split(m.qkv2(x2), m.num_heads)
        x12 = F.scaled_dot_product_attention(q1, k2, v2).transpose(1, 2).reshape(B, N, C)
        x21 = F.scaled_dot_product_attention(q2, k1, v1).transpose(1, 2).reshape(B, N, C)
        return m.proj(torch.cat((x12, x21), dim=-1))

    y, yr = m(x1, x2), ref(x1, x2)
    assert torch.allclose(y, yr, atol=1e-6)
    leaves = [x1, x2] + list(m.parameters())
    ga = torch.autograd.grad(y.square().sum(), leaves)
    gb = torch.autograd.grad(yr.square().sum(), leaves)
    assert all(torch.allclose(a, b, atol=1e-5) for a, b in zip(ga, gb))
