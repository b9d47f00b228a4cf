"""Fused glue around the mixer: adaLN modulate, gated residual (both with the token order folded into the row index),
residual-add + RMSNorm / LayerNorm (+ modulate), GatedMLP activation, CFG + Euler update.  Reference: `modulate`
dimsum/models_dim.py:34-35, the gated residuals :1510-1512 / :686-689, the transpose / flip copies :1498-1524, and the Triton
`rms_norm_fn` (mamba/mamba_ssm/ops/triton/layernorm.py:460).  One coalesced pass each; no permuted copy is materialised.
The raw kernels have no autograd; `modulate_fn`, `gate_residual_fn`, `add_rmsnorm_fn` and `gelu_mul_fn` at the end of the file
wrap them in autograd Functions for the recorded (training) pass: the backward is the same streaming kernels -- through the
inverse order table where the forward used one -- plus `dimsum_token_colsum`, `dimsum_gelu_mul_bwd`, `dimsum_add_rmsnorm_bwd`.
"""
import torch

from . import _lib

_DT = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def _rows(t, name):
    if t.dim() != 3 or t.stride(2) != 1 or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA (batch, seqlen, channels) tensor with channel stride 1")


def _vec(t, ref, name):
    if t.dim() != 2 or t.stride(1) != 1 or t.shape != (ref.shape[0], ref.shape[2]) or t.dtype not in _DT:
        raise RuntimeError(f"{name} must be (batch, channels) with stride(1) == 1")


def _idx(idx, L):
    if idx is None:
        return None
    if idx.dtype != torch.int32 or not idx.is_cuda or idx.numel() != L or not idx.is_contiguous():
        raise RuntimeError("idx must be a contiguous int32 CUDA tensor of length seqlen")
    return idx.data_ptr()


def modulate(x, shift, scale, idx=None, out_dtype=None):
    """out[b, l] = x[b, idx[l]] * (1 + scale[b]) + shift[b];  x, shift/scale and out may have different dtypes."""
    _rows(x, "x"); _vec(shift, x, "shift"); _vec(scale, x, "scale")
    if shift.stride(0) != scale.stride(0) or shift.dtype != scale.dtype:
        raise RuntimeError("shift and scale must share dtype and row stride (chunks of one adaLN output)")
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    with torch.cuda.device(x.device):
        p = _lib.RowwiseParams()
        p.batch, p.seqlen, p.channels = x.shape
        p.x_dtype, p.aux_dtype, p.dst_dtype = _DT[x.dtype], _DT[shift.dtype], _DT[out.dtype]
        p.x_batch_stride, p.x_token_stride = x.stride(0), x.stride(1)
        p.dst_batch_stride, p.dst_token_stride = out.stride(0), out.stride(1)
        p.vec_row_stride = shift.stride(0)
        p.x, p.shift, p.scale, p.dst = x.data_ptr(), shift.data_ptr(), scale.data_ptr(), out.data_ptr()
        p.idx = _idx(idx, x.shape[1])
        _lib.call("dimsum_modulate", p, torch.cuda.current_stream(x.device).cuda_stream)
    return out


def gate_residual(x, gate, m, idx=None, out_dtype=None):
    """out[b, l] = x[b, l] + gate[b] * m[b, idx[l]]; gate and m share a dtype that may differ from x's and the output's."""
    _rows(x, "x"); _rows(m, "m"); _vec(gate, x, "gate")
    if m.shape != x.shape or m.dtype != gate.dtype:
        raise RuntimeError("m must have the shape of x and the dtype of gate")
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype or x.dtype)
    with torch.cuda.device(x.device):
        p = _lib.RowwiseParams()
        p.batch, p.seqlen, p.channels = x.shape
        p.x_dtype, p.aux_dtype, p.dst_dtype = _DT[x.dtype], _DT[m.dtype], _DT[out.dtype]
        p.x_batch_stride, p.x_token_stride = x.stride(0), x.stride(1)
        p.m_batch_stride, p.m_token_stride = m.stride(0), m.stride(1)
        p.dst_batch_stride, p.dst_token_stride = out.stride(0), out.stride(1)
        p.vec_row_stride = gate.stride(0)
        p.x, p.m, p.gate, p.dst = x.data_ptr(), m.data_ptr(), gate.data_ptr(), out.data_ptr()
        p.idx = _idx(idx, x.shape[1])
        _lib.call("dimsum_gate_residual", p, torch.cuda.current_stream(x.device).cuda_stream)
    return out


def add_rmsnorm(x, residual, weight, eps, want_residual=True, out_dtype=None):
    """-> (y, res_out): res_out = x + residual in fp32, y = rmsnorm(res_out) * weight in `out_dtype` (default x.dtype)."""
    shape = x.shape
    x2 = x.reshape(-1, shape[-1])
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    rows, C = x2.shape
    if residual is not None:
        if residual.dtype != torch.float32 or residual.shape != shape:
            raise RuntimeError("add_rmsnorm: residual must be fp32 with the shape of x")
        residual = residual.contiguous()
    w = weight.float().contiguous()
    out_dtype = out_dtype or x.dtype
    y = torch.empty((rows, C), device=x.device, dtype=out_dtype)
    res_out = torch.empty((rows, C), device=x.device, dtype=torch.float32) if want_residual else None
    with torch.cuda.device(x.device):
        if out_dtype == x.dtype:
            p = _lib.RmsnormParams()
            p.rows, p.channels, p.dtype = rows, C, _DT[x.dtype]
            entry = "dimsum_add_rmsnorm"
        else:                       # same kernel through the entry point that takes the two dtypes separately (no modulate)
            p = _lib.NormModulateParams()
            p.rows, p.channels, p.rows_per_batch = rows, C, 1
            p.x_dtype, p.aux_dtype, p.y_dtype, p.norm_kind = _DT[x.dtype], _DT[torch.float32], _DT[out_dtype], 0
            p.shift = p.scale = None
            entry = "dimsum_norm_modulate"
        p.x_row_stride, p.y_row_stride = x2.stride(0), y.stride(0)
        p.x, p.weight, p.y = x2.data_ptr(), w.data_ptr(), y.data_ptr()
        p.residual = residual.data_ptr() if residual is not None else None
        p.res_out = res_out.data_ptr() if res_out is not None else None
        p.eps = eps
        _lib.call(entry, p, torch.cuda.current_stream(x.device).cuda_stream)
    return y.view(shape), (res_out.view(shape) if res_out is not None else None)


def norm_modulate(x, residual, weight, eps, shift, scale, layer_norm=False, out_dtype=None, want_residual=False):
    """-> (y, res_out).  h = x + residual (fp32); n = RMSNorm(h) * weight, or LayerNorm(h) without affine when
    `layer_norm`; y = n * (1 + scale[b]) + shift[b] in `out_dtype` (default x.dtype); res_out = h (fp32) on request.
    One pass for `hidden = hidden + x; modulate(norm_2(hidden), shift, scale)` (models_dim.py:1509-1512) and for
    `modulate(LayerNorm(x), shift, scale)` of the shared DiT block / final layer (models_dim.py:1079-1098)."""
    B, L, C = x.shape
    x2 = x.reshape(-1, C)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    if residual is not None:
        if residual.dtype != torch.float32 or residual.shape != x.shape:
            raise RuntimeError("norm_modulate: residual must be fp32 with the shape of x")
        residual = residual.contiguous()
    if shift.shape != (B, C) or scale.shape != (B, C) or shift.dtype != scale.dtype:
        raise RuntimeError("norm_modulate: shift / scale must be (batch, channels) of one dtype")
    if shift.stride(1) != 1 or scale.stride(1) != 1 or shift.stride(0) != scale.stride(0):
        shift, scale = shift.contiguous(), scale.contiguous()
    out_dtype = out_dtype or x.dtype
    y = torch.empty((B * L, C), device=x.device, dtype=out_dtype)
    res_out = torch.empty((B * L, C), device=x.device, dtype=torch.float32) if want_residual else None
    w = None if layer_norm else weight.float().contiguous()
    with torch.cuda.device(x.device):
        p = _lib.NormModulateParams()
        p.rows, p.channels, p.rows_per_batch = B * L, C, L
        p.x_dtype, p.aux_dtype, p.y_dtype, p.norm_kind = _DT[x.dtype], _DT[shift.dtype], _DT[out_dtype], int(layer_norm)
        p.x_row_stride, p.y_row_stride, p.vec_row_stride = x2.stride(0), y.stride(0), shift.stride(0)
        p.x, p.y = x2.data_ptr(), y.data_ptr()
        p.weight = w.data_ptr() if w is not None else None
        p.shift, p.scale = shift.data_ptr(), scale.data_ptr()
        p.residual = residual.data_ptr() if residual is not None else None
        p.res_out = res_out.data_ptr() if res_out is not None else None
        p.eps = eps
        _lib.call("dimsum_norm_modulate", p, torch.cuda.current_stream(x.device).cuda_stream)
    return y.view(B, L, C), (res_out.view(B, L, C) if res_out is not None else None)


def gelu_mul(x12):
    """GatedMLP activation: gelu_tanh(x12[..., :H]) * x12[..., H:] in one pass (dimsum/mlp.py:65-70)."""
    shape = x12.shape
    x2 = x12.reshape(-1, shape[-1])
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    rows, twoH = x2.shape
    H = twoH // 2
    y = torch.empty((rows, H), device=x12.device, dtype=x12.dtype)
    with torch.cuda.device(x12.device):
        p = _lib.GeluMulParams()
        p.rows, p.hidden, p.dtype = rows, H, _DT[x12.dtype]
        p.x_row_stride, p.y_row_stride = x2.stride(0), y.stride(0)
        p.x, p.y = x2.data_ptr(), y.data_ptr()
        _lib.call("dimsum_gelu_mul", p, torch.cuda.current_stream(x12.device).cuda_stream)
    return y.view(shape[:-1] + (H,))


def cfg_euler_step(x, model_out, cfg_scale, dt, out=None, v_out=None):
    """x_new = x + dt * (uncond + s * (cond - uncond)) on both halves of a CFG batch, one pass (models_dim.py:1886-1902 +
    integrators.py:98-111).  x (2n, C, H, W) fp32; model_out (2n, >= C, H, W), cond rows first; dt: fp32 CUDA scalar tensor."""
    n2, C, H, W = x.shape
    if (x.dtype != torch.float32 or not x.is_contiguous() or model_out.shape[0] != n2 or n2 % 2 or model_out.shape[1] < C
            or model_out.shape[2:] != x.shape[2:] or model_out.dtype not in _DT or model_out[0].numel() != model_out.stride(0)
            or dt.dtype != torch.float32 or not dt.is_cuda or dt.numel() != 1):
        raise RuntimeError("cfg_euler_step: x must be contiguous fp32 (2n, C, H, W), model_out (2n, >= C, H, W) with dense rows, dt an "
                           "fp32 CUDA scalar")
    out = torch.empty_like(x) if out is None else out
    with torch.cuda.device(x.device):
        p = _lib.CfgEulerParams()
        p.half_batch, p.channels, p.hw, p.out_dtype, p.out_row_stride = n2 // 2, C, H * W, _DT[model_out.dtype], model_out.stride(0)
        p.cfg_scale = float(cfg_scale)
        p.model_out, p.x, p.dt, p.x_new = model_out.data_ptr(), x.data_ptr(), dt.data_ptr(), out.data_ptr()
        if v_out is not None:
            if v_out.shape != x.shape or v_out.dtype != torch.float32 or not v_out.is_contiguous():
                raise RuntimeError("cfg_euler_step: v_out must be contiguous fp32 with the shape of x")
            p.v_out = v_out.data_ptr()
        _lib.call("dimsum_cfg_euler_step", p, torch.cuda.current_stream(x.device).cuda_stream)
    return out


# ---------------------------------------------------------------------------------------------------
# training: the same kernels under autograd
# ---------------------------------------------------------------------------------------------------
def token_colsum(g, x=None, want_sum_g=True, out_dtype=None, x_idx=None):
    """-> (sum_l g[b, l, :], sum_l g[b, l, :] * x[b, x_idx[l], :]) as (batch, channels) tensors (None where not requested);
    `x_idx` (int32 token table, default identity) pairs g's row l with x's row x_idx[l]."""
    _rows(g, "g")
    if x is not None:
        _rows(x, "x")
        if x.shape != g.shape:
            raise RuntimeError("token_colsum: x must have the shape of g")
    B, L, C = g.shape
    out_dtype = out_dtype or g.dtype
    sum_g = torch.empty((B, C), device=g.device, dtype=out_dtype) if want_sum_g else None
    sum_gx = torch.empty((B, C), device=g.device, dtype=out_dtype) if x is not None else None
    with torch.cuda.device(g.device):
        p = _lib.ColsumParams()
        p.batch, p.seqlen, p.channels = B, L, C
        p.g_dtype, p.x_dtype, p.out_dtype = _DT[g.dtype], _DT[x.dtype] if x is not None else 0, _DT[out_dtype]
        p.g_batch_stride, p.g_token_stride = g.stride(0), g.stride(1)
        if x is not None:
            p.x_batch_stride, p.x_token_stride = x.stride(0), x.stride(1)
        p.out_row_stride = C
        p.g, p.x = g.data_ptr(), x.data_ptr() if x is not None else None
        p.sum_g = sum_g.data_ptr() if sum_g is not None else None
        p.sum_gx = sum_gx.data_ptr() if sum_gx is not None else None
        p.x_idx = _idx(x_idx, L) if x is not None else None
        _lib.call("dimsum_token_colsum", p, torch.cuda.current_stream(g.device).cuda_stream)
    return sum_g, sum_gx


def _rows_ok(t):
    return t if t.stride(2) == 1 else t.contiguous()


class _ModulateFn(torch.autograd.Function):
    """y[l] = x[idx[l]] * (1 + scale[:, None]) + shift[:, None] (models_dim.py:34-35, with the scan order of
    models_dim.py:1498-1524 folded into the row index) in `out_dtype` (default: the promoted dtype of x and scale).
    `inv` is the inverse table of `idx`; both None = natural order."""

    @staticmethod
    def forward(ctx, x, shift, scale, idx, inv, out_dtype):
        x = _rows_ok(x)
        ctx.save_for_backward(x, scale)
        ctx.shift_dtype, ctx.idx, ctx.inv = shift.dtype, idx, inv
        return modulate(x, shift, scale, idx, out_dtype=out_dtype or torch.promote_types(x.dtype, scale.dtype))

    @staticmethod
    def backward(ctx, gy):
        x, scale = ctx.saved_tensors
        gy = _rows_ok(gy)
        gx = None
        if ctx.needs_input_grad[0]:             # gx[j] = gy[inv[j]] * (1 + scale): the same kernel through the inverse table
            zero = torch.zeros(scale.shape, device=scale.device, dtype=scale.dtype)
            sc = scale if zero.stride(0) == scale.stride(0) else scale.contiguous()
            gx = modulate(gy, zero, sc, ctx.inv, out_dtype=x.dtype)
        gshift, gscale = token_colsum(gy, x, want_sum_g=True, out_dtype=ctx.shift_dtype, x_idx=ctx.idx)
        return gx, gshift, gscale, None, None, None


class _GateResidualFn(torch.autograd.Function):
    """out[l] = x[l] + gate[:, None] * m[idx[l]] (models_dim.py:1510-1512; `idx` undoes the scan order, `inv` is its inverse)
    in `out_dtype` (default: the promoted dtype of x and gate)."""

    @staticmethod
    def forward(ctx, x, gate, m, idx, inv, out_dtype):
        x, m = _rows_ok(x), _rows_ok(m)
        ctx.save_for_backward(gate, m)
        ctx.x_dtype, ctx.idx, ctx.inv = x.dtype, idx, inv
        return gate_residual(x, gate, m, idx, out_dtype=out_dtype or torch.promote_types(x.dtype, gate.dtype))

    @staticmethod
    def backward(ctx, gy):
        gate, m = ctx.saved_tensors
        gy = _rows_ok(gy)
        gx = gy if gy.dtype == ctx.x_dtype else gy.to(ctx.x_dtype)
        gm = None
        if ctx.needs_input_grad[2]:
            # gm[k] = gate * gy[inv[k]] = gy * (1 + (gate - 1)) + 0.  `gate - 1` is formed in fp32: rounded to bf16 it would
            # quantise the effective gate to multiples of 2^-8 (a gate of 1e-3 would become 0); the kernel takes the aux
            # dtype independently of gy's
            gm1 = (gate.float() - 1).contiguous()
            gm = modulate(gy, torch.zeros_like(gm1), gm1, ctx.inv, out_dtype=m.dtype)
        _, ggate = token_colsum(gy, m, want_sum_g=False, out_dtype=gate.dtype, x_idx=ctx.idx)
        return gx, ggate, gm, None, None, None


class _GeluMulFn(torch.autograd.Function):
    """gelu_tanh(x12[..., :H]) * x12[..., H:] (mlp.py:65-70)."""

    @staticmethod
    def forward(ctx, x12):
        ctx.save_for_backward(x12)
        return gelu_mul(x12)

    @staticmethod
    def backward(ctx, gy):
        (x12,) = ctx.saved_tensors
        shape = x12.shape
        x2 = x12.reshape(-1, shape[-1])
        g2 = gy.reshape(-1, gy.shape[-1])
        if x2.stride(1) != 1:
            x2 = x2.contiguous()
        if g2.stride(1) != 1 or g2.dtype != x2.dtype:
            g2 = g2.to(x2.dtype).contiguous()
        rows, twoH = x2.shape
        dx = torch.empty((rows, twoH), device=x12.device, dtype=x12.dtype)
        with torch.cuda.device(x12.device):
            p = _lib.GeluMulBwdParams()
            p.rows, p.hidden, p.dtype = rows, twoH // 2, _DT[x12.dtype]
            p.x_row_stride, p.dy_row_stride, p.dx_row_stride = x2.stride(0), g2.stride(0), dx.stride(0)
            p.x, p.dy, p.dx = x2.data_ptr(), g2.data_ptr(), dx.data_ptr()
            _lib.call("dimsum_gelu_mul_bwd", p, torch.cuda.current_stream(x12.device).cuda_stream)
        return dx.view(shape)


class _AddRmsNormFn(torch.autograd.Function):
    """(y, h) = (rmsnorm(x + residual) * weight, x + residual in fp32); backward by `dimsum_add_rmsnorm_bwd`."""

    @staticmethod
    def forward(ctx, x, residual, weight, eps, out_dtype=None):
        y, h = add_rmsnorm(x, residual, weight, eps, want_residual=True, out_dtype=out_dtype)
        ctx.save_for_backward(h, weight)
        ctx.eps, ctx.x_dtype, ctx.has_res = eps, x.dtype, residual is not None
        return y, h

    @staticmethod
    def backward(ctx, gy, gh):
        h, weight = ctx.saved_tensors
        shape = h.shape
        C = shape[-1]
        h2 = h.reshape(-1, C)
        g2 = gy.reshape(-1, C)
        if g2.stride(1) != 1:
            g2 = g2.contiguous()
        if gh is not None:
            gh = gh.reshape(-1, C)
            gh = gh.float().contiguous() if (gh.dtype != torch.float32 or not gh.is_contiguous()) else gh
        rows = h2.shape[0]
        n_part = max(1, min((rows + 7) // 8, 2 * torch.cuda.get_device_properties(h.device).multi_processor_count))
        dx = torch.empty((rows, C), device=h.device, dtype=ctx.x_dtype)
        dres = torch.empty((rows, C), device=h.device, dtype=torch.float32) if ctx.has_res else None
        part = torch.empty((n_part, C), device=h.device, dtype=torch.float32)
        w = weight.float().contiguous()
        with torch.cuda.device(h.device):
            p = _lib.RmsnormBwdParams()
            p.rows, p.channels, p.n_partials = rows, C, n_part
            p.dy_dtype, p.dx_dtype = _DT[g2.dtype], _DT[dx.dtype]
            p.dy_row_stride, p.dx_row_stride = g2.stride(0), dx.stride(0)
            p.h, p.weight, p.dy = h2.data_ptr(), w.data_ptr(), g2.data_ptr()
            p.dres_in = gh.data_ptr() if gh is not None else None
            p.dx, p.dweight_partial = dx.data_ptr(), part.data_ptr()
            p.dres_out = dres.data_ptr() if dres is not None else None
            p.eps = ctx.eps
            _lib.call("dimsum_add_rmsnorm_bwd", p, torch.cuda.current_stream(h.device).cuda_stream)
        return dx.view(shape), (dres.view(shape) if dres is not None else None), part.sum(0).to(weight.dtype), None, None


def add_rmsnorm_fn(x, residual, weight, eps, out_dtype=None):
    """-> (y, h) with autograd: y = rmsnorm(x + residual) * weight in `out_dtype` (default x.dtype), h = x + residual in fp32."""
    return _AddRmsNormFn.apply(x, residual, weight, eps, out_dtype)


def modulate_fn(x, shift, scale, idx=None, inv=None, out_dtype=None):
    """Differentiable `modulate`; with a token table `idx` (and its inverse `inv`) the output row l is computed from row idx[l]."""
    if (idx is None) != (inv is None):
        raise RuntimeError("modulate_fn: idx and inv come together")
    return _ModulateFn.apply(x, shift, scale, idx, inv, out_dtype)


def gate_residual_fn(x, gate, m, idx=None, inv=None, out_dtype=None):
    """Differentiable `gate_residual`; with a token table `idx` (and its inverse `inv`) row l adds gate * m[idx[l]]."""
    if (idx is None) != (inv is None):
        raise RuntimeError("gate_residual_fn: idx and inv come together")
    return _GateResidualFn.apply(x, gate, m if m.dtype == gate.dtype else m.to(gate.dtype), idx, inv, out_dtype)


def gelu_mul_fn(x12):
    return _GeluMulFn.apply(x12)
