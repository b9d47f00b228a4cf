"""TEST INFRASTRUCTURE -- functional CPU restatement of the reference DiM forward on top of `oracle.ref_ops`.

`dim_forward_oracle(sd, x, t, y)` evaluates the released DiMSUM wiring (dimsum/models_dim.py:1796-1884 `DiM.forward`,
:1055-1117 `DiMBlockCombined.forward`, :1447-1524 `DiMBlockRaw.forward`, :606-705 `WaveDiMBlock.forward`,
:1532-1554 `DiTBlock`, :205-220 `FinalLayer`, attention_fusion.py:61-84, mlp.py:65-70) directly from a reference
state dict, in the reference's own order of operations: materialised transpose / flip / local_scan copies, the 2-level DWT
as the closed form of `ref_ops.wavelet_packet_oracle`, and the Mamba slow path (mamba_simple.py:658-700) on
`selective_scan_oracle` / `causal_conv1d_oracle`.  Pinned against the unmodified reference model by
tests/test_oracle_golden.py::test_model_oracle_matches_reference (tests/golden/model_*.npz).

Used as (a) the model-level checker for the GPU tests and smoke(), (b) the timed CPU baseline / `--impl reference`
arm of bench.py.  Never imported by the product package.  Every function is device-agnostic torch code: fed CUDA tensors
it is the "same-device oracle" of SURVEY.md section 8c (reference ops as ordinary CUDA tensor ops, GEMMs identical on both
sides), which is what tests/test_model_fullsize_gpu.py uses for the DiM-L/2-sized comparisons.
"""
import math

import torch
import torch.nn.functional as F

from . import orders, ref_ops


def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _lin(sd, prefix, x):
    return F.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _adaln(sd, prefix, c, n):
    return _lin(sd, prefix + ".adaLN_modulation.1", F.silu(c)).chunk(n, dim=1)


def _gated_mlp(sd, prefix, x):
    x1, x2 = _lin(sd, prefix + ".w12", x).chunk(2, dim=-1)
    return _lin(sd, prefix + ".w3", F.gelu(x1, approximate="tanh") * x2)


def _heads(sd, prefix, x, n_heads):
    B, N, C = x.shape
    return _lin(sd, prefix, x).reshape(B, N, 3, n_heads, C // n_heads).permute(2, 0, 3, 1, 4).unbind(0)


def _mamba_slow_path(sd, prefix, h):
    """CondMamba.forward, use_fast_path=False branch (mamba_simple.py:658-700); cond_proj is dead."""
    Bsz, L, _ = h.shape
    A = -torch.exp(sd[prefix + ".A_log"].float())
    Dm, N = A.shape
    rank = sd[prefix + ".dt_proj.weight"].shape[1]
    xz = (sd[prefix + ".in_proj.weight"] @ h.reshape(Bsz * L, -1).t()).view(2 * Dm, Bsz, L).transpose(0, 1)
    x, z = xz[:, :Dm], xz[:, Dm:]
    x = ref_ops.causal_conv1d_oracle(x, sd[prefix + ".conv1d.weight"].reshape(Dm, -1), sd[prefix + ".conv1d.bias"], "silu")
    x_dbl = F.linear(x.transpose(1, 2).reshape(Bsz * L, Dm), sd[prefix + ".x_proj.weight"])
    dt = (sd[prefix + ".dt_proj.weight"] @ x_dbl[:, :rank].t()).view(Dm, Bsz, L).transpose(0, 1)
    Bm = x_dbl[:, rank:rank + N].view(Bsz, L, N).transpose(1, 2).contiguous()
    Cm = x_dbl[:, rank + N:].view(Bsz, L, N).transpose(1, 2).contiguous()
    y = ref_ops.selective_scan_oracle(x, dt, A, Bm, Cm, sd[prefix + ".D"].float(), z=z,
                                      delta_bias=sd[prefix + ".dt_proj.bias"].float(), delta_softplus=True)
    return _lin(sd, prefix + ".out_proj", y.transpose(1, 2))


def _rms(x, w, eps=1e-5):
    return ref_ops.rms_norm_oracle(x, w, eps=eps)


def dim_forward_oracle(sd, x, t, y, depth=None, attn_every=4, in_channels=4, patch=2, attn_heads=16):
    hidden = sd["pos_embed"].shape[-1]
    depth = depth if depth is not None else 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    # embedders (models_dim.py:1808-1814)
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    args = t[:, None].float() * freqs[None]
    temb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    c = _lin(sd, "t_embedder.mlp.2", F.silu(_lin(sd, "t_embedder.mlp.0", temb))) + sd["y_embedder.embedding_table.weight"][y]
    h = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=patch).flatten(2).transpose(1, 2)
    h = h + sd["pos_embed"]
    L = h.shape[1]
    grid = math.isqrt(L)
    residual = None
    for i in range(depth):
        p = f"blocks.{i}"
        reverse, transpose = i % 2 > 0, i % 4 >= 2
        h, residual = ref_ops.rms_norm_oracle(h, sd[p + ".norm.weight"], residual=residual, eps=1e-5, prenorm=True)
        x1, x2 = h.chunk(2, dim=2)
        # spatial branch: materialised order copies exactly like DiMBlockRaw.forward
        seq_np = orders.implicit_spatial_order(grid, transpose, reverse)
        seq = torch.from_numpy(seq_np).to(h.device)
        inv = torch.from_numpy(orders.invert(seq_np)).to(h.device)
        s = x1[:, seq]
        sh, sc, g = _adaln(sd, p + ".spatial_mamba", c, 3)
        s = s + g.unsqueeze(1) * _mamba_slow_path(sd, p + ".spatial_mamba.mixer", _modulate(s, sh, sc))
        x1 = s[:, inv]
        # frequency branch: dwt -> window scan -> mamba -> inverse
        f = ref_ops.window_scan_oracle(ref_ops.wavelet_packet_oracle(x2), grid // 4, column_first=reverse)
        sh, sc, g = _adaln(sd, p + ".freq_mamba", c, 3)
        f = f + g.unsqueeze(1) * _mamba_slow_path(sd, p + ".freq_mamba.mixer", _modulate(f, sh, sc))
        x2 = ref_ops.wavelet_packet_inverse_oracle(ref_ops.window_unscan_oracle(f, grid // 4, column_first=reverse))
        # cross-attention fusion, 8 heads (attention_fusion.py:61-84)
        q1, k1, v1 = _heads(sd, p + ".proj.qkv1", x1, 8)
        q2, k2, v2 = _heads(sd, p + ".proj.qkv2", x2, 8)
        Bsz, _, C = x1.shape
        x12 = F.scaled_dot_product_attention(q1, k2, v2).transpose(1, 2).reshape(Bsz, L, C)
        x21 = F.scaled_dot_product_attention(q2, k1, v1).transpose(1, 2).reshape(Bsz, L, C)
        h = h + _lin(sd, p + ".proj.proj", torch.cat((x12, x21), dim=-1))
        sh, sc, g = _adaln(sd, p, c, 3)
        h = h + g.unsqueeze(1) * _gated_mlp(sd, p + ".mlp", _modulate(_rms(h, sd[p + ".norm_2.weight"]), sh, sc))
        if attn_every > 0 and (i + 1) % attn_every == 0:
            s1, c1, g1, s2, c2, g2 = _adaln(sd, "attn_block", c, 6)
            a = _modulate(F.layer_norm(h, (hidden,), eps=1e-6), s1, c1)
            q, k, v = _heads(sd, "attn_block.attn.qkv", a, attn_heads)
            a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(h.shape)
            h = h + g1.unsqueeze(1) * _lin(sd, "attn_block.attn.proj", a)
            h = h + g2.unsqueeze(1) * _gated_mlp(sd, "attn_block.mlp", _modulate(F.layer_norm(h, (hidden,), eps=1e-6), s2, c2))
    sh, sc = _adaln(sd, "final_layer", c, 2)
    out = _lin(sd, "final_layer.linear", _modulate(F.layer_norm(h, (hidden,), eps=1e-6), sh, sc))
    n_out = out.shape[-1] // (patch * patch)
    out = out.reshape(out.shape[0], grid, grid, patch, patch, n_out)
    return torch.einsum("nhwpqc->nchpwq", out).reshape(out.shape[0], n_out, grid * patch, grid * patch)


def dim_forward_with_cfg_oracle(sd, x, t, y, cfg_scale, **kw):
    """models_dim.py:1886-1902."""
    half = x[: len(x) // 2]
    out = dim_forward_oracle(sd, torch.cat([half, half], dim=0), t, y, **kw)
    cond, uncond = torch.split(out, len(out) // 2, dim=0)
    g = uncond + cfg_scale * (cond - uncond)
    return torch.cat([g, g], dim=0)


def euler_sample_oracle(sd, z, y, cfg_scale, num_steps=250, null_class=1000, **kw):
    """Fixed-grid Euler on linspace(0, 1, num_steps) of the velocity ODE (transport.py:181-183, integrators.py:98-111;
    torchdiffeq 0.2.3 `euler` on a fixed grid is x_{i+1} = x_i + (t_{i+1} - t_i) f(t_i, x_i)) with CFG batching as in
    sample_ddp.py:168-178."""
    x = torch.cat([z, z], dim=0)
    yy = torch.cat([y, torch.full_like(y, null_class)], dim=0)
    ts = torch.linspace(0, 1, num_steps, device=z.device)
    for i in range(num_steps - 1):
        v = dim_forward_with_cfg_oracle(sd, x, torch.ones(x.shape[0], device=z.device) * ts[i], yy, cfg_scale, **kw)
        x = x + (ts[i + 1] - ts[i]) * v
    return x[: len(z)]
