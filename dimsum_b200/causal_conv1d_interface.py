"""Autograd front end of the depthwise causal convolution on the B200 kernels.

Public name and argument meaning follow the reference (`causal_conv1d_fn`, causal-conv1d/causal_conv1d/
causal_conv1d_interface.py:37-46): x (batch, dim, seqlen), weight (dim, width), optional bias (dim,), activation in
{None, "silu", "swish"}; the result has the shape and dtype of x.  Everything is executed by
`dimsum_causal_conv1d_fwd` / `dimsum_causal_conv1d_bwd` through `causal_conv1d_cuda`; there is no PyTorch path.
"""
import torch

from . import causal_conv1d_cuda as _native

_ACTIVATIONS = {None: False, "silu": True, "swish": True}


def _sequence_major(t):
    """The kernels want the sequence axis innermost; a channel-last tensor is passed through so that the native layer
    can reject it explicitly, anything else irregular is compacted first (as the reference does)."""
    return t if t.stride(2) == 1 or t.stride(1) == 1 else t.contiguous()


class CausalConv1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias=None, activation=None):
        if activation not in _ACTIVATIONS:
            raise NotImplementedError("activation must be None, silu, or swish")
        use_silu = _ACTIVATIONS[activation]
        x = _sequence_major(x)
        if bias is not None:
            bias = bias.contiguous()
        out = _native.causal_conv1d_fwd(x, weight, bias, use_silu)
        ctx.use_silu = use_silu
        ctx.save_for_backward(x, weight, bias)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight, bias = ctx.saved_tensors
        grad_x, grad_w, grad_b = _native.causal_conv1d_bwd(x, weight, bias, _sequence_major(grad_out), None, ctx.use_silu)
        return grad_x, grad_w, (grad_b if bias is not None else None), None


def causal_conv1d_fn(x, weight, bias=None, activation=None):
    """out[b, d, l] = act(bias[d] + sum_w weight[d, w] * x[b, d, l - (width - 1) + w]) with zero history."""
    return CausalConv1dFn.apply(x, weight, bias, activation)
