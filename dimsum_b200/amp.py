"""bf16 autocast training without the per-step weight casts.

Under `torch.autocast` every fp32 weight is cast to bf16 once per forward and every bf16 weight gradient is cast back to
fp32 once per backward: ~800 tiny kernels and two passes over the 0.5 G parameters per DiM-L/2 step (reference loop:
dimsum/train.py:302-321 under `torch.autocast`).  `Bf16Shadows` keeps a bf16 copy of every `nn.Linear` weight / bias next
to its fp32 master, refreshed after the optimizer step by ONE multi-tensor copy; `linear()` / `weight_times_rows_t()` run the
GEMMs on the shadows and produce the weight gradient directly in fp32 (bf16 x bf16 -> fp32 GEMM output, no cast pass) and
the bias gradient with this repo's column-sum kernel.  Values are the ones autocast computes: the same bf16 roundings of
weights and activations, fp32 accumulation; the weight gradient skips autocast's intermediate rounding to bf16.

Without shadows (the default) everything here falls through to the plain PyTorch expressions.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F


def shadow_of(param):
    """bf16 shadow of a master parameter when one is attached and autocast to bf16 is on for its device -- else None.  A shadow
    that is older than its master (the optimizer stepped without `Bf16Shadows.refresh()`) is brought up to date on the spot,
    which costs the cast autocast would have done."""
    if param is None:
        return None
    sh = getattr(param, "_dimsum_bf16", None)
    if sh is None:
        return None
    dev = param.device.type
    if (not torch.is_autocast_enabled(dev) or torch.get_autocast_dtype(dev) != sh.dtype or sh.device != param.device
            or param.dtype != torch.float32 or sh.shape != param.shape):
        return None                                   # no autocast, or the master was moved / recast / resized since
    if param._dimsum_bf16_version != param._version:
        sh.copy_(param.detach())
        param._dimsum_bf16_version = param._version
    return sh


class Bf16Shadows:
    """Attach bf16 shadows to every `nn.Linear` weight and bias of `model`; call `refresh()` after each optimizer step (inside
    the captured graph when the step is replayed as a CUDA graph)."""

    def __init__(self, model, dtype=torch.bfloat16):
        self.masters, self.shadows = [], []
        seen = set()
        for mod in model.modules():
            if not isinstance(mod, nn.Linear):
                continue
            for prm in (mod.weight, mod.bias):
                if prm is None or id(prm) in seen or prm.dtype != torch.float32:
                    continue
                seen.add(id(prm))
                prm._dimsum_bf16 = torch.empty(prm.shape, device=prm.device, dtype=dtype)
                prm._dimsum_bf16_version = -1
                self.masters.append(prm)
                self.shadows.append(prm._dimsum_bf16)
        self.refresh()

    @torch.no_grad()
    def refresh(self):
        if self.masters:
            torch._foreach_copy_(self.shadows, [p.detach() for p in self.masters])
            for p in self.masters:
                p._dimsum_bf16_version = p._version

    def detach(self):
        for p in self.masters:
            del p._dimsum_bf16, p._dimsum_bf16_version
        self.masters, self.shadows = [], []


def mm_wgrad(a, b, master_dtype):
    """a @ b for a weight gradient: written in the master's dtype by the GEMM itself when the operands are 16-bit."""
    if a.is_cuda and master_dtype == torch.float32 and a.dtype in (torch.bfloat16, torch.float16) and a.dtype == b.dtype:
        return torch.mm(a, b, out_dtype=torch.float32)
    return a @ b


def bias_grad(g2, master_dtype):
    """Column sum of a (rows, channels) gradient in fp32: this repo's two-stage column-sum kernel (deterministic) when the
    shape allows it, else `sum(0)`."""
    rows, C = g2.shape
    if g2.is_cuda and g2.stride(1) == 1 and C % 4 == 0 and g2.dtype in (torch.float32, torch.bfloat16, torch.float16):
        from . import fused
        groups = next((n for n in (32, 16, 8, 4, 2) if rows % n == 0 and rows // n >= 8), 1)
        if g2.stride(0) == C and groups > 1:
            part, _ = fused.token_colsum(g2.view(groups, rows // groups, C), out_dtype=torch.float32)
            return part.sum(0).to(master_dtype)
    return g2.sum(0, dtype=torch.float32).to(master_dtype)


class _ShadowLinearFn(torch.autograd.Function):
    """F.linear(x, weight, bias) computed on the bf16 shadows (w16, b16); gradients go to the fp32 masters."""

    @staticmethod
    def forward(ctx, x, weight, bias, w16, b16):
        x16 = x if x.dtype == w16.dtype else x.to(w16.dtype)
        ctx.save_for_backward(x16, w16)
        ctx.x_shape, ctx.has_bias = x.shape, bias is not None
        ctx.w_dtype = weight.dtype
        return F.linear(x16, w16, b16)

    @staticmethod
    def backward(ctx, gy):
        x16, w16 = ctx.saved_tensors
        g2 = gy.reshape(-1, gy.shape[-1])
        if g2.dtype != w16.dtype:
            g2 = g2.to(w16.dtype)
        gx = (g2 @ w16).view(ctx.x_shape) if ctx.needs_input_grad[0] else None
        gw = mm_wgrad(g2.t(), x16.reshape(-1, x16.shape[-1]), ctx.w_dtype) if ctx.needs_input_grad[1] else None
        gb = bias_grad(g2, ctx.w_dtype) if ctx.has_bias and ctx.needs_input_grad[2] else None
        return gx, gw, gb, None, None


def linear(x, weight, bias=None):
    """F.linear; on the bf16 shadows when `weight` has one (see `Bf16Shadows`) and autocast to bf16 is on."""
    w16 = shadow_of(weight)
    if w16 is None:
        return F.linear(x, weight, bias)
    return _ShadowLinearFn.apply(x, weight, bias, w16, shadow_of(bias))


class _ShadowWeightTimesRowsTFn(torch.autograd.Function):
    """weight @ rows.t() -> (out_features, n_rows) on the bf16 shadow (the transposed in_proj of the mixers)."""

    @staticmethod
    def forward(ctx, rows, weight, w16):
        r16 = rows if rows.dtype == w16.dtype else rows.to(w16.dtype)
        ctx.save_for_backward(r16, w16)
        ctx.w_dtype = weight.dtype
        return w16 @ r16.t()

    @staticmethod
    def backward(ctx, g):
        r16, w16 = ctx.saved_tensors
        if g.dtype != w16.dtype:
            g = g.to(w16.dtype)
        grows = g.t() @ w16 if ctx.needs_input_grad[0] else None
        gw = mm_wgrad(g, r16, ctx.w_dtype) if ctx.needs_input_grad[1] else None
        return grows, gw, None


def weight_times_rows_t(weight, rows):
    """weight (E, D) @ rows (n, D)^T -> (E, n); shadow-aware like `linear`."""
    w16 = shadow_of(weight)
    if w16 is None:
        return weight @ rows.t()
    return _ShadowWeightTimesRowsTFn.apply(rows, weight, w16)
