"""DiM backbone in the released DiMSUM wiring, on the B200 hot-path kernels.

Reference: dimsum/models_dim.py -- `DiM` (:1557-1930), `DiMBlockCombined` (:974-1117), `DiMBlockRaw` (:1402-1530),
`WaveDiMBlock` (:505-708), `DiTBlock` (:1532-1554), `FinalLayer` (:205-220), embedders (:129-203); fusion and MLP from
dimsum/attention_fusion.py:9-84 and dimsum/mlp.py:49-70.

Only what the released checkpoints use is built: block_type="combined", cond_mamba, rms_norm, fused_add_norm,
learnable_pe, pe_type="ape", shared attention block every k layers, 2 wavelet levels.  Module and parameter names
follow the reference one-to-one, so `load_state_dict` of a reference checkpoint works with strict=True.

Hot-path differences from the reference (same results, fewer HBM passes):
  * spatial branch: the transpose / flip orders (models_dim.py:1498-1524) are not materialised; the order table is
    handed to the mixer, whose conv and scan kernels read / write through it;
  * frequency branch: `_dwt_fast` + `local_scan` is one kernel, `local_reverse` + `_idwt_fast` is one kernel;
  * the dead `cond_proj` GEMM and its (B, d_inner, L) buffer are skipped.
Dense layers (in/x/dt/out projections, attention, MLP) are cuBLAS / SDPA calls, as in the reference.
"""
import math
import os
from functools import partial

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import amp, fused
from . import scanning_orders as so
from .attention import attention, attention_supported
from .mamba_simple import CondMamba
from .wavelet import wavelet_packet, wavelet_packet_inverse


def modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def _fused_ok(x):
    """The one-pass CUDA glue kernels have no autograd: use them exactly when nothing is being recorded."""
    return x.is_cuda and not torch.is_grad_enabled()


_GLUE_DTYPES = (torch.float32, torch.bfloat16, torch.float16)


def _train_fused_ok(x, *others):
    """Recorded (training) pass on CUDA: the same glue kernels under autograd (fused.modulate_fn & co) unless
    DIMSUM_TRAIN_FUSED=0 asks for the plain PyTorch expressions."""
    return (x.is_cuda and x.dim() == 3 and x.shape[-1] % 8 == 0 and os.environ.get("DIMSUM_TRAIN_FUSED", "1") != "0"
            and all(t.dtype in _GLUE_DTYPES for t in (x,) + others))


def _amp_dtype():
    """Under autocast the consumer of a glue kernel's output is a low-precision GEMM: emit its input dtype directly instead
    of an fp32 tensor that autocast would re-read and cast (same single rounding, one pass less forward and backward)."""
    return torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled() else None


def _mod(x, shift, scale, idx=None, inv=None):
    """modulate(x[idx], shift, scale); every caller feeds the result to a GEMM.  `inv` is the inverse table of `idx`."""
    if _fused_ok(x):
        return fused.modulate(x, shift, scale, idx, out_dtype=_amp_dtype())
    if _train_fused_ok(x, shift, scale):
        return fused.modulate_fn(x, shift, scale, idx, inv, out_dtype=_amp_dtype())
    assert idx is None
    return modulate(x, shift, scale)


def _norm_mod(norm, x, shift, scale, residual=None):
    """modulate(norm(x + residual), shift, scale) -> (modulated, x + residual); one kernel when nothing is recorded."""
    is_rms = isinstance(norm, RMSNorm)
    if (_fused_ok(x) and x.dim() == 3 and x.dtype in (torch.float32, torch.bfloat16, torch.float16)
            and (residual is None or residual.dtype == torch.float32)
            and (is_rms or (isinstance(norm, nn.LayerNorm) and norm.weight is None and norm.bias is None))):
        y, h = fused.norm_modulate(x, residual, norm.weight if is_rms else None, norm.eps, shift, scale,
                                   layer_norm=not is_rms, out_dtype=_amp_dtype(), want_residual=residual is not None)
        return y, (h if residual is not None else x)
    if (is_rms and residual is not None and residual.dtype == torch.float32 and _train_fused_ok(x, shift, scale)
            and x.shape[-1] <= 1024 and norm.weight.dtype == torch.float32 and torch.is_grad_enabled()):
        # recorded pass: the residual add rides on the norm kernel (fp32 sum out, fp32 normalised rows), then modulate
        n, h = fused.add_rmsnorm_fn(x, residual, norm.weight, norm.eps, out_dtype=torch.float32)
        return _mod(n, shift, scale), h
    h = x if residual is None else x + residual
    return _mod(norm(h), shift, scale), h


def _gated(x, gate, m, idx=None, inv=None, feeds_gemm=False):
    """x + gate * m[idx]; `inv` is the inverse table of `idx`.  A result that only feeds low-precision GEMMs under autocast is
    emitted in their dtype (no cast pass)."""
    if _fused_ok(x):
        if m.dtype != gate.dtype:
            m = m.to(gate.dtype)
        return fused.gate_residual(x, gate, m, idx, out_dtype=_amp_dtype() if feeds_gemm else None)
    if _train_fused_ok(x, gate, m) and m.shape == x.shape:
        return fused.gate_residual_fn(x, gate, m, idx, inv, out_dtype=_amp_dtype() if feeds_gemm else None)
    assert idx is None
    return x + gate.unsqueeze(1) * m


class _Cond:
    """The conditioning vector c = t_emb + y_emb together with SiLU(c): every adaLN head of the model is Linear(SiLU(c)) on
    the same c (models_dim.py:1079, 1509, 1546), so DiM.forward evaluates the activation (and, under autocast, its cast to the
    GEMM dtype) once instead of once per head -- 53 heads in DiM-L/2."""
    __slots__ = ("raw", "act")

    def __init__(self, c):
        self.raw = c
        act = F.silu(c)
        dt = _amp_dtype()
        self.act = act.to(dt) if dt is not None and c.is_cuda else act


def _raw(c):
    return c.raw if isinstance(c, _Cond) else c


def _ada(head, c):
    """adaLN head (nn.Sequential(SiLU, Linear)) on a conditioning vector or on a `_Cond`."""
    if isinstance(c, _Cond):
        return head[1](c.act) if isinstance(head[0], nn.SiLU) else head(c.raw)
    return head(c)


class Linear(nn.Linear):
    """nn.Linear whose GEMMs run on the bf16 shadow of the weight when one is attached (`amp.Bf16Shadows`, training)."""

    def forward(self, x):
        return amp.linear(x, self.weight, self.bias)


class RMSNorm(nn.Module):
    """Residual-add + RMSNorm with fp32 residual (reference: Triton `rms_norm_fn`, layernorm.py:460; maths of
    `rms_norm_ref`, layernorm.py:32-47)."""

    def __init__(self, hidden_size, eps=1e-5, device=None, dtype=None):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(hidden_size, device=device, dtype=dtype))
        self.register_parameter("bias", None)

    def forward(self, x, residual=None, prenorm=False, residual_in_fp32=False):
        if _fused_ok(x) and (residual is None or residual.dtype == torch.float32) and (residual_in_fp32 or not prenorm):
            y, res = fused.add_rmsnorm(x, residual, self.weight, self.eps, want_residual=prenorm)
            return (y, res) if prenorm else y
        if (_train_fused_ok(x) and x.shape[-1] <= 1024 and x.shape[-1] % 8 == 0 and self.weight.dtype == torch.float32
                and (residual is None or residual.dtype == torch.float32) and (residual_in_fp32 or not prenorm)):
            y, res = fused.add_rmsnorm_fn(x, residual, self.weight, self.eps)
            return (y, res) if prenorm else y
        io_dtype = x.dtype
        xf = x.float()
        if residual is not None:
            xf = xf + residual.float()
        y = (xf * torch.rsqrt(xf.square().mean(-1, keepdim=True) + self.eps) * self.weight.float()).to(io_dtype)
        if not prenorm:
            return y
        return y, (xf if residual_in_fp32 else xf.to(io_dtype))


class TimestepEmbedder(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):
        super().__init__()
        self.mlp = nn.Sequential(Linear(frequency_embedding_size, hidden_size), nn.SiLU(), Linear(hidden_size, hidden_size))
        self.frequency_embedding_size = frequency_embedding_size

    @staticmethod
    def timestep_embedding(t, dim, max_period=10000):
        half = dim // 2
        freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
        args = t[:, None].float() * freqs[None]
        emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
        if dim % 2:
            emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
        return emb

    def forward(self, t):
        return self.mlp(self.timestep_embedding(t, self.frequency_embedding_size))


class LabelEmbedder(nn.Module):
    def __init__(self, num_classes, hidden_size, dropout_prob):
        super().__init__()
        self.in_channels = num_classes + int(dropout_prob > 0)
        self.embedding_table = nn.Embedding(self.in_channels, hidden_size)
        self.num_classes, self.dropout_prob = num_classes, dropout_prob

    def forward(self, labels, train, force_drop_ids=None):
        if (train and self.dropout_prob > 0) or force_drop_ids is not None:
            drop = torch.rand(labels.shape[0], device=labels.device) < self.dropout_prob if force_drop_ids is None \
                else force_drop_ids == 1
            labels = torch.where(drop, self.num_classes, labels)
        return self.embedding_table(labels)

    def get_in_channels(self):
        return self.in_channels


class PatchEmbed(nn.Module):
    def __init__(self, img_size, patch_size, in_chans, embed_dim):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class _QkvSplitFn(torch.autograd.Function):
    """(B, N, 3*H*D) -> q, k, v as (B, H, N, D) views, like `.view(B, N, 3, H, D).permute(2, 0, 3, 1, 4).unbind(0)`
    (attention_fusion.py:63-70).  Autograd's own backward of that chain stacks the three gradients into a (3, B, H, N, D)
    tensor and then permutes it into a second copy -- two slow strided passes per attention; here each gradient is written
    once, straight into its slot of the (B, N, 3, H, D) gradient of the projection."""

    @staticmethod
    def forward(ctx, qkv, num_heads):
        B, N, C3 = qkv.shape
        ctx.shape = (B, N, num_heads, C3 // 3 // num_heads)
        return qkv.view(B, N, 3, num_heads, -1).permute(2, 0, 3, 1, 4).unbind(0)

    @staticmethod
    def backward(ctx, *grads):
        B, N, H, D = ctx.shape
        ref = next(g for g in grads if g is not None)
        out = torch.empty((B, N, 3, H, D), device=ref.device, dtype=ref.dtype)
        for i, g in enumerate(grads):
            if g is None:
                out[:, :, i].zero_()
            else:
                out[:, :, i].copy_(g.transpose(1, 2))
        return out.view(B, N, 3 * H * D), None


def _split_qkv(qkv, num_heads):
    if qkv.requires_grad and torch.is_grad_enabled():
        return _QkvSplitFn.apply(qkv, num_heads)
    B, N, C3 = qkv.shape
    return qkv.view(B, N, 3, num_heads, C3 // 3 // num_heads).permute(2, 0, 3, 1, 4).unbind(0)


class Attention(nn.Module):
    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = Linear(dim, dim)

    def forward(self, x):
        B, N, C = x.shape
        q, k, v = _split_qkv(self.qkv(x), self.num_heads)
        if attention_supported(q, k, v):          # fp32 sampling with TF32 allowed: the TMA + tcgen05 kernel of this repo
            return self.proj(attention(q, k, v))  # already (B, N, C): no transpose copy
        return self.proj(F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, N, C))


class GatedMLP(nn.Module):
    def __init__(self, in_features, hidden_features, act_layer, bias=True):
        super().__init__()
        self.w12 = Linear(in_features, 2 * hidden_features, bias=bias)
        self.w3 = Linear(hidden_features, in_features, bias=bias)
        self.act_layer = act_layer()
        # the one-pass kernels hard-wire tanh-GELU (mlp.py:65-70 with the released act_layer); any other activation runs
        # the plain PyTorch expression in eval and in training alike
        self._tanh_gelu = isinstance(self.act_layer, nn.GELU) and self.act_layer.approximate == "tanh"

    def forward(self, x):
        x12 = self.w12(x)
        if self._tanh_gelu and _fused_ok(x12) and x12.dtype in (torch.float32, torch.bfloat16, torch.float16):
            return self.w3(fused.gelu_mul(x12))
        if self._tanh_gelu and _train_fused_ok(x12):
            return self.w3(fused.gelu_mul_fn(x12))
        x1, x2 = x12.chunk(2, dim=-1)
        return self.w3(self.act_layer(x1) * x2)


class CrossAttentionFusion(nn.Module):
    """Each half queries the other half's keys/values (attention_fusion.py:61-84, swap_k=False)."""

    def __init__(self, dim, num_heads=8, qkv_bias=True):
        super().__init__()
        self.num_heads, self.head_dim = num_heads, dim // 2 // num_heads
        self.qkv1 = Linear(dim // 2, dim // 2 * 3, bias=qkv_bias)
        self.qkv2 = Linear(dim // 2, dim // 2 * 3, bias=qkv_bias)
        self.proj = Linear(dim, dim)

    def forward(self, x1, x2):
        B, N, C = x1.shape
        q1, k1, v1 = _split_qkv(self.qkv1(x1), self.num_heads)
        q2, k2, v2 = _split_qkv(self.qkv2(x2), self.num_heads)
        if attention_supported(q1, k2, v2) and attention_supported(q2, k1, v1):
            # both cross attentions write their halves of the (B, N, 2C) input of `proj` directly: no transpose, no cat
            both = torch.empty((B, N, 2 * C), device=x1.device, dtype=q1.dtype)
            attention(q1, k2, v2, out=both[:, :, :C].view(B, N, self.num_heads, self.head_dim))
            attention(q2, k1, v1, out=both[:, :, C:].view(B, N, self.num_heads, self.head_dim))
            return self.proj(both)
        # cat((x12, x21), -1) of the head-merged outputs == the two (B, N, H, D) outputs stacked along the head axis: one
        # pass instead of two transpose copies and a cat
        o12 = F.scaled_dot_product_attention(q1, k2, v2).transpose(1, 2)
        o21 = F.scaled_dot_product_attention(q2, k1, v1).transpose(1, 2)
        return self.proj(torch.cat((o12, o21), dim=2).view(B, N, 2 * C))


def _block_order(mixer, order, inv, device):
    """Sequence order a block has to present to its mixer: the block's implicit order composed with the mixer's scan table
    (sequence position k reads token order[table[k]]), and its inverse; (None, None) when there is nothing to reorder."""
    table = mixer.table_order(device) if hasattr(mixer, "table_order") else None
    if table is None:
        return order, inv
    # keyed on the table's storage and version too: load_state_dict after a first forward overwrites zigzag_paths in place
    key = (device, None if order is None else order.data_ptr(), table.data_ptr(), table._version)
    cache = mixer.__dict__.setdefault("_block_order_cache", {})
    if key not in cache:
        total = table if order is None else order[table.long()].contiguous()
        cache[key] = (total, mixer.inverse_order(total))
    return cache[key]


def _order_buffer(table):
    return torch.from_numpy(np.ascontiguousarray(table).astype(np.int32))


class DiMBlockRaw(nn.Module):
    """Spatial Mamba branch: x + gate * mixer(modulate(x)) scanned in {row, column} x {forward, reversed} order."""

    def __init__(self, dim, mixer_cls, c_dim, grid, reverse=False, transpose=False):
        super().__init__()
        self.reverse, self.transpose = reverse, transpose
        self.mixer = mixer_cls(dim)
        self.norm = nn.Identity()
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), Linear(c_dim, 3 * dim, bias=True))
        order = so.implicit_order(grid, transpose, reverse) if (reverse or transpose) else None
        self.register_buffer("_order", _order_buffer(order) if order is not None else None, persistent=False)
        self.register_buffer("_inv", _order_buffer(so.reverse_permut_np(order)) if order is not None else None,
                             persistent=False)

    def forward(self, x, c):
        shift, scale, gate = _ada(self.adaLN_modulation, c).chunk(3, dim=1)
        if _fused_ok(x) or _train_fused_ok(x, shift, scale, gate):
            # the scan order (implicit transpose / flip, or the mixer's zigma / sweep / jpeg table) rides on the row index
            # of the two glue kernels -- and of their backward kernels when the pass is recorded: no permuted copy, no
            # extra pass, and the mixer runs gather-free
            order, inv = _block_order(self.mixer, self._order, self._inv, x.device)
            m = self.mixer(_mod(x, shift, scale, order, inv), _raw(c), pre_ordered=True)
            return _gated(x, gate, m, inv, order, feeds_gemm=True)
        return x + gate.unsqueeze(1) * self.mixer(modulate(x, shift, scale), _raw(c), order=self._order)


class WaveDiMBlock(nn.Module):
    """Frequency branch: wavelet packet -> window scan -> Mamba -> inverse (models_dim.py:606-705, no_ffn=True)."""

    def __init__(self, dim, mixer_cls, c_dim, grid, column_first=False, num_wavelet_lv=2):
        super().__init__()
        if num_wavelet_lv != 2:
            raise NotImplementedError("only the released 2-level wavelet packet is implemented")
        self.mixer = mixer_cls(dim)
        self.norm = nn.Identity()
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), Linear(c_dim, 3 * dim, bias=True))
        # Haar filter buffers exist in reference checkpoints (wavelet_layer.py:71-89,95-114); kept so strict loading works
        s = 0.5
        self.dwt = nn.Module()
        for name, sg in (("w_ll", (1, 1, 1, 1)), ("w_lh", (1, 1, -1, -1)), ("w_hl", (1, -1, 1, -1)), ("w_hh", (1, -1, -1, 1))):
            self.dwt.register_buffer(name, torch.tensor(sg, dtype=torch.float32).view(1, 1, 2, 2) * s)
        self.idwt = nn.Module()
        self.idwt.register_buffer("filters", torch.stack([getattr(self.dwt, n)[0] for n in ("w_ll", "w_lh", "w_hl", "w_hh")]))
        seq_to_token = so.window_order(grid, grid // 4, column_first)
        self.register_buffer("_pos", _order_buffer(so.reverse_permut_np(seq_to_token)), persistent=False)

    def forward(self, x, c):
        h = wavelet_packet(x, self._pos)                                 # _dwt_fast + local_scan
        shift, scale, gate = _ada(self.adaLN_modulation, c).chunk(3, dim=1)
        if _fused_ok(h) or _train_fused_ok(h, shift, scale, gate):
            order, inv = _block_order(self.mixer, None, None, h.device)      # the mixer's own table, if its scan type has one
            h = _gated(h, gate, self.mixer(_mod(h, shift, scale, order, inv), _raw(c), pre_ordered=True), inv, order, feeds_gemm=True)
        else:
            h = _gated(h, gate, self.mixer(_mod(h, shift, scale), _raw(c)), feeds_gemm=True)
        return wavelet_packet_inverse(h, self._pos)                      # local_reverse + _idwt_fast


class DiMBlockCombined(nn.Module):
    def __init__(self, dim, mixer_cls, grid, reverse=False, transpose=False, eps=1e-5):
        super().__init__()
        self.norm = RMSNorm(dim, eps=eps)
        self.spatial_mamba = DiMBlockRaw(dim // 2, mixer_cls, c_dim=dim, grid=grid, reverse=reverse, transpose=transpose)
        self.freq_mamba = WaveDiMBlock(dim // 2, mixer_cls, c_dim=dim, grid=grid, column_first=reverse)
        self.proj = CrossAttentionFusion(dim, num_heads=8, qkv_bias=True)
        self.norm_2 = RMSNorm(dim, eps=eps)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), Linear(dim, 3 * dim, bias=True))
        self.mlp = GatedMLP(dim, int(dim * 4), act_layer=lambda: nn.GELU(approximate="tanh"))

    def forward(self, hidden_states, residual, c):
        hidden_states, residual = self.norm(hidden_states, residual=residual, prenorm=True, residual_in_fp32=True)
        x1, x2 = hidden_states.chunk(2, dim=2)
        x = self.proj(self.spatial_mamba(x1, c), self.freq_mamba(x2, c))
        shift, scale, gate = _ada(self.adaLN_modulation, c).chunk(3, dim=1)
        m, hidden_states = _norm_mod(self.norm_2, x, shift, scale, residual=hidden_states)      # hidden + x, norm_2, modulate
        hidden_states = _gated(hidden_states, gate, self.mlp(m))
        return hidden_states, residual


class DiTBlock(nn.Module):
    def __init__(self, hidden_size, num_heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = Attention(hidden_size, num_heads=num_heads, qkv_bias=True)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = GatedMLP(hidden_size, int(hidden_size * 4), act_layer=lambda: nn.GELU(approximate="tanh"))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), Linear(hidden_size, 6 * hidden_size, bias=True))

    def forward(self, x, c):
        s1, sc1, g1, s2, sc2, g2 = _ada(self.adaLN_modulation, c).chunk(6, dim=1)
        x = _gated(x, g1, self.attn(_norm_mod(self.norm1, x, s1, sc1)[0]))
        return _gated(x, g2, self.mlp(_norm_mod(self.norm2, x, s2, sc2)[0]))


class FinalLayer(nn.Module):
    def __init__(self, hidden_size, patch_size, out_channels):
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), Linear(hidden_size, 2 * hidden_size, bias=True))

    def forward(self, x, c):
        shift, scale = _ada(self.adaLN_modulation, c).chunk(2, dim=1)
        return self.linear(_norm_mod(self.norm_final, x, shift, scale)[0])


def get_2d_sincos_pos_embed(embed_dim, grid_size):
    def one_d(dim, pos):
        omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    gw, gh = np.meshgrid(np.arange(grid_size, dtype=np.float32), np.arange(grid_size, dtype=np.float32))
    return np.concatenate([one_d(embed_dim // 2, gw), one_d(embed_dim // 2, gh)], axis=1)


class DiM(nn.Module):
    def __init__(self, img_resolution=32, patch_size=2, in_channels=4, hidden_size=1024, depth=16, label_dropout=0.1,
                 num_classes=1000, learn_sigma=False, ssm_cfg=None, rms_norm=True, residual_in_fp32=True, fused_add_norm=True,
                 scan_type="none", pe_type="ape", block_type="combined", cond_mamba=True, scanning_continuity=False,
                 learnable_pe=True, drop_path=0.0, use_final_norm=False, use_attn_every_k_layers=4, use_gated_mlp=True,
                 **unused):
        super().__init__()
        unsupported = dict(block_type=(block_type, "combined"), pe_type=(pe_type, "ape"), rms_norm=(rms_norm, True),
                           fused_add_norm=(fused_add_norm, True), cond_mamba=(cond_mamba, True),
                           scanning_continuity=(scanning_continuity, False), use_final_norm=(use_final_norm, False),
                           use_gated_mlp=(use_gated_mlp, True), drop_path=(drop_path, 0.0))
        for k, (got, want) in unsupported.items():
            if got != want:
                raise NotImplementedError(f"DiM: {k}={got!r} is not part of the released DiMSUM configuration ({want!r})")
        self.depth, self.learn_sigma, self.in_channels = depth, learn_sigma, in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size, self.num_classes = patch_size, num_classes
        self.use_attn_every_k_layers = use_attn_every_k_layers
        self.x_embedder = PatchEmbed(img_resolution, patch_size, in_channels, hidden_size)
        self.t_embedder = TimestepEmbedder(hidden_size)
        self.y_embedder = LabelEmbedder(num_classes, hidden_size, label_dropout)
        num_patches = self.x_embedder.num_patches
        grid = int(math.isqrt(num_patches))
        self.pos_embed = nn.Parameter(torch.zeros(1, num_patches, hidden_size), requires_grad=learnable_pe)
        mixer_kwargs = dict(ssm_cfg or {})
        mixer_kwargs.pop("use_fast_path", None)
        if scan_type.startswith(("zigma", "sweep", "jpeg")):        # gen_paths, models_dim.py:1640-1658
            kind, n_paths = scan_type.split("_")[0], int(scan_type.split("_")[1])
            paths = so.SCAN_ZOO[kind](grid)[:n_paths]
            mixer_kwargs["zigzag_paths"] = torch.from_numpy(np.stack(paths * depth))
            mixer_kwargs["zigzag_paths_reverse"] = torch.from_numpy(np.stack([so.reverse_permut_np(p) for p in paths] * depth))
        elif scan_type != "none":
            raise NotImplementedError(f"DiM: scan_type={scan_type!r}")
        self.blocks = nn.ModuleList([
            DiMBlockCombined(hidden_size,
                             partial(CondMamba, layer_idx=i, scan_type=scan_type, d_cond=hidden_size, **mixer_kwargs),
                             grid=grid, reverse=(scan_type == "none") and (i % 2 > 0),
                             transpose=(scan_type == "none") and (i % 4 >= 2))
            for i in range(depth)])
        if use_attn_every_k_layers > 0:
            self.attn_block = DiTBlock(hidden_size, 16)
        self.final_layer = FinalLayer(hidden_size, patch_size, self.out_channels)
        self.initialize_weights()

    def initialize_weights(self):
        """Same scheme as models_dim.py:1744-1786 (adaLN-zero, zero final layer, GPT-2 style out_proj rescale)."""
        grid = int(math.isqrt(self.x_embedder.num_patches))
        self.pos_embed.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(self.pos_embed.shape[-1], grid)).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.zeros_(self.x_embedder.proj.bias)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for block in self.blocks:
            nn.init.zeros_(block.adaLN_modulation[-1].weight)
            nn.init.zeros_(block.adaLN_modulation[-1].bias)
        for m in (self.final_layer.adaLN_modulation[-1], self.final_layer.linear):
            nn.init.zeros_(m.weight)
            nn.init.zeros_(m.bias)
        for module in self.modules():
            if isinstance(module, nn.Linear) and module.bias is not None and not getattr(module.bias, "_no_reinit", False):
                nn.init.zeros_(module.bias)
            elif isinstance(module, nn.Embedding):
                nn.init.normal_(module.weight, std=0.02)
        for name, p in self.named_parameters():
            if name.endswith("out_proj.weight") or name.endswith("fc2.weight"):
                nn.init.kaiming_uniform_(p, a=math.sqrt(5))
                with torch.no_grad():
                    p /= math.sqrt(self.depth)

    def unpatchify(self, x):
        c, p = self.out_channels, self.patch_size
        h = w = int(math.isqrt(x.shape[1]))
        x = x.reshape(x.shape[0], h, w, p, p, c)
        return torch.einsum("nhwpqc->nchpwq", x).reshape(x.shape[0], c, h * p, w * p)

    def forward(self, x, t, y=None, inference_params=None, **kwargs):
        """x (N, C, H, W) latents, t (N,) times, y (N,) labels -> (N, C_out, H, W); models_dim.py:1796-1884."""
        if y is None:
            y = torch.full((x.size(0),), self.y_embedder.get_in_channels() - 1, dtype=torch.long, device=x.device)
        c = _Cond(self.t_embedder(t) + self.y_embedder(y, self.training))
        x = self.x_embedder(x) + self.pos_embed
        residual = None
        for idx, block in enumerate(self.blocks):
            x, residual = block(x, residual, c)
            if self.use_attn_every_k_layers > 0 and (idx + 1) % self.use_attn_every_k_layers == 0:
                x = self.attn_block(x, c)
        return self.unpatchify(self.final_layer(x, c))

    def forward_with_cfg(self, x, t, y=None, inference_params=None, cfg_scale=1.0, **kwargs):
        """Classifier-free guidance on all channels of the first in_channels (models_dim.py:1886-1902)."""
        half = x[: len(x) // 2]
        out = self.forward(torch.cat([half, half], dim=0), t, y)
        eps, rest = out[:, : self.in_channels], out[:, self.in_channels:]
        cond, uncond = torch.split(eps, len(eps) // 2, dim=0)
        half_eps = uncond + cfg_scale * (cond - uncond)
        return torch.cat([torch.cat([half_eps, half_eps], dim=0), rest], dim=1)


def _zoo(depth, hidden_size, patch_size):
    return lambda **kwargs: DiM(depth=depth, hidden_size=hidden_size, patch_size=patch_size, **kwargs)


# models_dim.py:2163-2236
DiM_models = {
    "DiM-XL/2": _zoo(24, 1152, 2),
    "DiM-L/2": _zoo(16, 1024, 2),
    "DiM-L/2-v1": _zoo(20, 1024, 2),
    "DiM-B/2": _zoo(12, 768, 2),
    "DiM-L/4": _zoo(16, 1024, 4),
    "DiM-L/4-v1": _zoo(20, 1024, 4),
}
