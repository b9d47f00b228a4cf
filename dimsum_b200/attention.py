"""Softmax attention of the fusion / DiT blocks on the TMA + tcgen05 kernel (`dimsum_attention_fwd`).

Reference call sites: `F.scaled_dot_product_attention(q, k, v)` in dimsum/attention_fusion.py:61-84 and the timm `Attention`
of the shared DiTBlock (dimsum/models_dim.py:1532-1554).  fp32 tensors, TF32 tensor-core products with fp32 accumulation and
an fp32 softmax -- the precision class cuBLAS uses for the surrounding GEMMs when `torch.backends.cuda.matmul.allow_tf32` is
on, which is the condition for taking this path.  One CTA per (batch, head, 128-query tile) walks the keys in blocks of 128
with the online softmax, two CTAs per SM (256 tokens at 256px, 1024 at 512px).  Forward only: recorded
(training) passes and 16-bit autocast keep the library SDPA (a library GPU kernel, not a fallback of this repo's hot path).
"""
import os

import torch

from . import _lib


def attention_supported(q, k, v):
    if os.environ.get("DIMSUM_ATTENTION", "1") == "0" or torch.is_grad_enabled() or not torch.backends.cuda.matmul.allow_tf32:
        return False
    if not (q.is_cuda and q.dtype == k.dtype == v.dtype == torch.float32 and q.dim() == k.dim() == v.dim() == 4):
        return False
    B, H, Nq, d = q.shape
    Nk = k.shape[2]
    if d != 64 or k.shape != (B, H, Nk, d) or v.shape != k.shape or Nk % 64 or B > 65535 or H > 65535:
        return False
    for t in (q, k, v):
        if t.stride(3) != 1 or t.data_ptr() % 16 or any(s <= 0 or s % 4 for s in t.stride()[:3]):
            return False
    return True


def attention(q, k, v, out=None):
    """q, k, v: (batch, heads, tokens, 64) fp32 views -> (batch, tokens_q, heads * 64), the layout the output projection
    reads.  `out`: optional (batch, tokens_q, heads, 64) view to write into (e.g. one half of a wider buffer)."""
    B, H, Nq, d = q.shape
    Nk = k.shape[2]
    if out is None:
        out = torch.empty((B, Nq, H, d), device=q.device, dtype=q.dtype)
    elif out.shape != (B, Nq, H, d) or out.dtype != q.dtype or out.stride(3) != 1 or out.data_ptr() % 16 \
            or any(s <= 0 or s % 4 for s in out.stride()[:3]):
        raise RuntimeError("attention: out must be a (batch, tokens, heads, 64) fp32 view with 16-byte aligned rows")
    with torch.cuda.device(q.device):
        p = _lib.AttentionParams()
        p.batch, p.heads, p.seqlen_q, p.seqlen_k, p.head_dim, p.dtype = B, H, Nq, Nk, d, _lib.DTYPE_F32
        p.q_batch_stride, p.q_head_stride, p.q_token_stride = q.stride(0), q.stride(1), q.stride(2)
        p.k_batch_stride, p.k_head_stride, p.k_token_stride = k.stride(0), k.stride(1), k.stride(2)
        p.v_batch_stride, p.v_head_stride, p.v_token_stride = v.stride(0), v.stride(1), v.stride(2)
        p.out_batch_stride, p.out_token_stride, p.out_head_stride = out.stride(0), out.stride(1), out.stride(2)
        p.scale = d ** -0.5
        p.q, p.k, p.v, p.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
        _lib.call("dimsum_attention_fwd", p, torch.cuda.current_stream(q.device).cuda_stream)
    return out.view(B, Nq, H * d) if out.is_contiguous() else out
