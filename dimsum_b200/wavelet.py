"""Fused 2-level Haar wavelet-packet split / merge (reference: dimsum/wavelet_layer.py + WaveDiMBlock._dwt_fast /
_idwt_fast, dimsum/models_dim.py:572-604), optionally fused with the window scan order (local_scan / local_reverse,
models_dim.py:662,701).

    wavelet_packet(x, pos=None)          == local_scan(_dwt_fast(x))          when pos = inverse window order
    wavelet_packet_inverse(y, pos=None)  == _idwt_fast(local_reverse(y))

x, y: (batch, grid*grid, channels) token-major, channels % 16 == 0, grid % 4 == 0.  `pos[token]` is the sequence
position at which the transformed token is stored (forward) / found (inverse); int32 CUDA tensor or None.
Both directions are one kernel launch; each is the other's gradient up to the 1/16 scale.
"""
import math

import torch

from . import _lib

_DT = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def _run(x, pos, inverse, scale):
    if x.dim() != 3 or x.dtype not in _DT or not x.is_cuda:
        raise RuntimeError("wavelet_packet: x must be a CUDA (batch, tokens, channels) float tensor")
    if x.stride(2) != 1:
        x = x.contiguous()
    B, L, C = x.shape
    grid = math.isqrt(L)
    if grid * grid != L or grid % 4 or C % 16:
        raise RuntimeError("wavelet_packet: tokens must form a square grid divisible by 4 and channels % 16 == 0")
    if pos is not None and (pos.dtype != torch.int32 or not pos.is_cuda or pos.numel() != L or not pos.is_contiguous()):
        raise RuntimeError("wavelet_packet: pos must be a contiguous int32 CUDA tensor with one entry per token")
    out = torch.empty((B, L, C), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        p = _lib.WaveletParams()
        p.batch, p.grid, p.channels, p.dtype = B, grid, C, _DT[x.dtype]
        p.src_batch_stride, p.src_token_stride = x.stride(0), x.stride(1)
        p.dst_batch_stride, p.dst_token_stride = out.stride(0), out.stride(1)
        p.src, p.dst = x.data_ptr(), out.data_ptr()
        p.pos = pos.data_ptr() if pos is not None else None
        p.scale = scale
        _lib.call("dimsum_wavelet_packet_inv" if inverse else "dimsum_wavelet_packet_fwd", p,
                  torch.cuda.current_stream(x.device).cuda_stream)
    return out


class _WaveletPacketFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pos, inverse):
        ctx.pos, ctx.inverse = pos, inverse
        return _run(x, pos, inverse, 1.0 if inverse else 1.0 / 16.0)

    @staticmethod
    def backward(ctx, g):
        # d(fwd) = inverse butterfly * 1/16 ; d(inv) = forward butterfly * 1   (wavelet_layer.py:22-33,50-65)
        return _run(g, ctx.pos, not ctx.inverse, 1.0 / 16.0 if not ctx.inverse else 1.0), None, None


def wavelet_packet(x, pos=None):
    return _WaveletPacketFn.apply(x, pos, False)


def wavelet_packet_inverse(x, pos=None):
    return _WaveletPacketFn.apply(x, pos, True)
