"""GPU parity of the TMA + tcgen05 attention kernel (`dimsum_attention_fwd`) against an fp64 softmax attention on the same
values (reference call: F.scaled_dot_product_attention, dimsum/attention_fusion.py:61-84, models_dim.py:1532-1554).  The
kernel multiplies in TF32 (10-bit mantissa, fp32 accumulation, fp32 softmax), the precision class of the cuBLAS GEMMs around it
under allow_tf32; the tolerance below is that class's (2e-3 of the output's max norm), and the model-level test checks that
a DiM forward with it stays within the same bound of the forward with the library SDPA."""
import pytest
import torch

from golden_io import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _tf32_on():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    yield
    torch.backends.cuda.matmul.allow_tf32 = old


def _ref(q, k, v):
    s = (q.double() @ k.double().transpose(-1, -2)) * q.shape[-1] ** -0.5
    return (torch.softmax(s, dim=-1) @ v.double()).transpose(1, 2).reshape(q.shape[0], q.shape[2], -1)


@pytest.mark.parametrize("B,H,Nq,Nk", [(3, 8, 256, 256), (2, 16, 256, 256), (2, 8, 128, 128), (1, 4, 64, 64), (2, 8, 256, 128),
                                      (2, 2, 200, 192), (1, 1, 384, 256), (2, 4, 1024, 1024), (1, 2, 300, 576), (1, 3, 128, 320)])
def test_attention_matches_fp64_softmax_attention(B, H, Nq, Nk):
    from dimsum_b200.attention import attention, attention_supported
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Nq)
    # q / k / v as the slices of fused qkv projections (the model's layout: (B, N, 3, H, 64) permuted), large logits included
    scale = 1.5 if B % 2 else 1.0                      # sharp and flat softmax rows
    qkv_q = torch.randn(B, Nq, 3, H, 64, generator=g, device="cuda") * scale
    qkv_k = torch.randn(B, Nk, 3, H, 64, generator=g, device="cuda") * scale
    q = qkv_q.permute(2, 0, 3, 1, 4)[0]
    k, v = qkv_k.permute(2, 0, 3, 1, 4)[1], qkv_k.permute(2, 0, 3, 1, 4)[2]
    with torch.no_grad():
        assert attention_supported(q, k, v)
        got = attention(q, k, v)
        assert got.shape == (B, Nq, H * 64)
        assert rel_err(got, _ref(q, k, v).float()) <= 2e-3, rel_err(got, _ref(q, k, v).float())
        # writing into one half of a wider buffer (CrossAttentionFusion: no cat copy)
        both = torch.full((B, Nq, 2 * H * 64), 9.0, device="cuda")
        attention(q, k, v, out=both[:, :, H * 64:].view(B, Nq, H, 64))
        assert torch.equal(both[:, :, H * 64:], got) and bool((both[:, :, :H * 64] == 9.0).all())
        # contiguous (B, H, N, 64) inputs
        got2 = attention(q.contiguous(), k.contiguous(), v.contiguous())
        assert torch.equal(got2, got)


def test_attention_is_only_taken_when_its_conditions_hold():
    from dimsum_b200.attention import attention_supported
    q = torch.randn(2, 8, 256, 64, device="cuda")
    with torch.no_grad():
        assert attention_supported(q, q, q)
        assert not attention_supported(q.half(), q.half(), q.half())                     # 16-bit keeps the library flash kernel
        assert attention_supported(q, torch.randn(2, 8, 1024, 64, device="cuda"), torch.randn(2, 8, 1024, 64, device="cuda"))
        assert not attention_supported(q, torch.randn(2, 8, 1000, 64, device="cuda"), torch.randn(2, 8, 1000, 64, device="cuda"))
        assert not attention_supported(q[..., :32], q[..., :32], q[..., :32])            # head_dim 64 only
        torch.backends.cuda.matmul.allow_tf32 = False
        assert not attention_supported(q, q, q)                                          # TF32 disabled: library SDPA
        torch.backends.cuda.matmul.allow_tf32 = True
    assert not attention_supported(q, q, q)                                              # recording gradients: library SDPA


def test_model_forward_with_tcgen05_attention_matches_library_sdpa(monkeypatch):
    """DiM (hidden 1024 -> heads of 64, 256 tokens, 5 blocks incl. the shared DiT block) with this repo's attention vs the same
    forward with F.scaled_dot_product_attention, both with TF32 GEMMs."""
    from dimsum_b200.models_dim import DiM
    torch.manual_seed(0)
    m = DiM(img_resolution=32, in_channels=4, hidden_size=1024, depth=4, num_classes=10, use_attn_every_k_layers=2).cuda().eval()
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "adaLN_modulation" in n or n.startswith("final_layer.linear"):
                p.normal_(0, 0.02)
    x = torch.randn(3, 4, 32, 32, device="cuda")
    t = torch.rand(3, device="cuda")
    y = torch.randint(0, 10, (3,), device="cuda")
    with torch.no_grad():
        from dimsum_b200 import _lib
        before = _lib.launch_count()
        ours = m(x, t, y)
        n_ours = _lib.launch_count() - before
        monkeypatch.setenv("DIMSUM_ATTENTION", "0")
        before = _lib.launch_count()
        lib = m(x, t, y)
        n_lib = _lib.launch_count() - before
    assert n_ours == n_lib + 4 * 2 + 2                      # two cross attentions per block + the DiT block twice
    assert rel_err(ours, lib) <= 2e-3, rel_err(ours, lib)
