"""Drop-in for the reference's pybind module `causal_conv1d_cuda` (causal-conv1d/csrc/causal_conv1d.cpp:571-577).

Same function names, argument order, checks and return values.  `causal_conv1d_fwd_cond` keeps the reference's
observable contract: the conditioning tensor is only the output buffer and is overwritten
(causal_conv1d.cpp:326).  The channel-last layout and the decode-time `causal_conv1d_update` are outside this
hot path and raise NotImplementedError (no fallback).
"""
import torch

from . import _lib

_DT = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _checks(x, weight, bias_):
    _check(x.dtype in _DT, "causal_conv1d: input must be float32, float16 or bfloat16")
    _check(weight.dtype in _DT, "causal_conv1d: weight must be float32, float16 or bfloat16")
    _check(x.is_cuda and weight.is_cuda, "causal_conv1d: x and weight must be CUDA tensors")
    _check(x.dim() == 3 and weight.dim() == 2, "causal_conv1d: x must be (batch, dim, seqlen), weight (dim, width)")
    batch, dim, seqlen = x.shape
    width = weight.shape[1]
    _check(weight.shape[0] == dim, "causal_conv1d: weight shape mismatch")
    _check(2 <= width <= 4, "causal_conv1d only supports width between 2 and 4")
    if x.stride(2) != 1:
        if x.stride(1) == 1:
            raise NotImplementedError("causal_conv1d: channel-last layout is not implemented in the B200 kernels")
        raise RuntimeError("causal_conv1d: x must have stride(2) == 1")
    if bias_ is not None:
        _check(bias_.dtype == weight.dtype and bias_.is_cuda and bias_.stride(-1) == 1 and tuple(bias_.shape) == (dim,),
               "causal_conv1d: bias must be a contiguous (dim,) CUDA tensor of the weight dtype")
    return batch, dim, seqlen, width


def _fwd_into(x, weight, bias_, silu_activation, out, perm=None):
    batch, dim, seqlen, width = _checks(x, weight, bias_)
    with torch.cuda.device(x.device):
        p = _lib.ConvFwdParams()
        p.batch, p.dim, p.seqlen, p.width = batch, dim, seqlen, width
        p.io_dtype, p.w_dtype, p.silu = _DT[x.dtype], _DT[weight.dtype], int(bool(silu_activation))
        p.x_batch_stride, p.x_d_stride = x.stride(0), x.stride(1)
        p.out_batch_stride, p.out_d_stride = out.stride(0), out.stride(1)
        p.w_d_stride, p.w_width_stride = weight.stride(0), weight.stride(1)
        p.x, p.weight, p.out = x.data_ptr(), weight.data_ptr(), out.data_ptr()
        p.bias = bias_.data_ptr() if bias_ is not None else None
        if perm is not None:
            _check(perm.dtype == torch.int32 and perm.is_cuda and perm.is_contiguous() and perm.numel() == seqlen,
                   "causal_conv1d: perm must be a contiguous int32 CUDA tensor of length seqlen")
            p.perm = perm.data_ptr()
        _lib.call("dimsum_causal_conv1d_fwd", p, _stream(x))
    return out


def causal_conv1d_fwd(x, weight, bias_, silu_activation, *, perm=None):
    """causal_conv1d.cpp:221-281.  out = empty_like(x)."""
    _checks(x, weight, bias_)
    out = torch.empty(x.shape, device=x.device, dtype=x.dtype)
    return _fwd_into(x, weight, bias_, silu_activation, out, perm)


def causal_conv1d_fwd_cond(x, weight, bias_, silu_activation, init_x):
    """causal_conv1d.cpp:283-347: `at::Tensor out = init_x;` -- writes the result into init_x and returns it."""
    _checks(x, weight, bias_)
    if init_x is None:
        return causal_conv1d_fwd(x, weight, bias_, silu_activation)
    _check(init_x.dtype == x.dtype and init_x.is_cuda and init_x.shape == x.shape and init_x.stride(2) == 1,
           "causal_conv1d_fwd_cond: init_x must match x in dtype and shape")
    return _fwd_into(x, weight, bias_, silu_activation, init_x)


def conv_xproj_supported(x, weight, x_proj_weight, out=None, precise=False):
    """Shapes / layouts the fused conv + x_proj kernel takes (anything else runs the two separate steps)."""
    if precise and x.dtype == torch.float32 and x_proj_weight.shape[0] > 64:
        return False                                  # hi + lo weight tiles of more than 64 rows do not fit the shared-memory ring
    if x.dtype not in _DT or x_proj_weight.dtype != x.dtype or not x.is_cuda or x.dim() != 3 or x.stride(2) != 1:
        return False
    es = x.element_size()
    vec, kc = 16 // es, 128 // es
    if out is not None and not (out.dtype == x.dtype and out.is_cuda and out.shape == x.shape and out.stride(2) == 1
                                and out.stride(0) % vec == 0 and out.stride(1) % vec == 0 and out.data_ptr() % 16 == 0):
        return False
    n_out, dim = x_proj_weight.shape
    return (dim == x.shape[1] and dim % kc == 0 and x.shape[2] % vec == 0 and 8 <= n_out <= 128 and n_out % 8 == 0
            and x_proj_weight.stride(1) == 1 and x_proj_weight.stride(0) % vec == 0 and x_proj_weight.data_ptr() % 16 == 0
            and x.stride(0) % vec == 0 and x.stride(1) % vec == 0 and x.data_ptr() % 16 == 0
            and weight.dtype == torch.float32 and weight.dim() == 2 and weight.shape[1] == 4 and weight.is_contiguous()
            and weight.data_ptr() % 16 == 0 and 0 < x.shape[0] <= 65535)


def conv_xproj_fwd(x, weight, bias_, x_proj_weight, precise=False, out=None, split=None):
    """u = silu(causal_conv1d(x)) and x_dbl = x_proj(u) in ONE tcgen05 kernel (selective_scan_interface.py:836-840 fused).

    -> (u, x_dbl) with x_dbl (batch, n_out, seqlen) channel-major, or, with `split` = dt_rank (a multiple of 8),
    -> (u, dt, bc): dt (split, batch * seqlen) -- the right operand of the dt_proj GEMM, so delta = (W_dt @ dt) comes out in
    the (dim, batch, seqlen) layout of selective_scan_interface.py:841 -- and bc (batch, n_out - split, seqlen), whose halves
    are B and C in the layout the scan reads (no rearrange copies).
    `out` (the reference's `init_states` buffer, causal_conv1d.cpp:326) receives u when given.  `precise` selects the
    3xTF32 split for fp32 I/O (fp32-grade results, used when TF32 matmuls are disabled)."""
    batch, dim, seqlen, width = _checks(x, weight, bias_)
    _check(conv_xproj_supported(x, weight, x_proj_weight, precise=precise),
           "conv_xproj_fwd: unsupported shape or layout (see conv_xproj_supported)")
    if bias_ is None:
        bias_ = torch.zeros(dim, device=x.device, dtype=weight.dtype)
    precise = bool(precise) and x.dtype == torch.float32
    # low part of the 3xTF32 split of the weight: w - (w with the 13 low mantissa bits cleared), exact in fp32
    xw_lo = (x_proj_weight - (x_proj_weight.view(torch.int32) & -8192).view(torch.float32)) if precise else None
    vec = 16 // x.element_size()
    if out is not None:
        _check(out.dtype == x.dtype and out.is_cuda and out.shape == x.shape and out.stride(2) == 1
               and out.stride(0) % vec == 0 and out.stride(1) % vec == 0 and out.data_ptr() % 16 == 0,
               "conv_xproj_fwd: out must match x in dtype and shape with 16-byte aligned rows")
    n_out = x_proj_weight.shape[0]
    if split is not None:
        _check(0 < split < n_out and split % 8 == 0, "conv_xproj_fwd: split must be a multiple of 8 inside (0, n_out)")
    with torch.cuda.device(x.device):
        u = out if out is not None else torch.empty(x.shape, device=x.device, dtype=x.dtype)
        p = _lib.ConvXprojParams()
        p.batch, p.dim, p.seqlen, p.width, p.n_out = batch, dim, seqlen, width, n_out
        p.io_dtype, p.w_dtype, p.precision = _DT[x.dtype], _DT[weight.dtype], int(precise)
        p.x_proj_weight_lo = xw_lo.data_ptr() if precise else None
        p.x_batch_stride, p.x_d_stride = x.stride(0), x.stride(1)
        p.u_batch_stride, p.u_d_stride = u.stride(0), u.stride(1)
        if split is None:
            x_dbl = torch.empty((batch, n_out, seqlen), device=x.device, dtype=x.dtype)
            p.x_dbl_batch_stride, p.x_dbl_row_stride = x_dbl.stride(0), x_dbl.stride(1)
            p.x_dbl = x_dbl.data_ptr()
            result = (u, x_dbl)
        else:
            dt = torch.empty((split, batch * seqlen), device=x.device, dtype=x.dtype)
            bc = torch.empty((batch, n_out - split, seqlen), device=x.device, dtype=x.dtype)
            p.split_rows = split
            p.x_dbl_batch_stride, p.x_dbl_row_stride = seqlen, batch * seqlen
            p.tail_batch_stride, p.tail_row_stride = bc.stride(0), bc.stride(1)
            p.x_dbl, p.x_dbl_tail = dt.data_ptr(), bc.data_ptr()
            result = (u, dt, bc)
        p.w_d_stride, p.w_width_stride, p.xw_row_stride = weight.stride(0), weight.stride(1), x_proj_weight.stride(0)
        p.x, p.conv_weight, p.x_proj_weight = x.data_ptr(), weight.data_ptr(), x_proj_weight.data_ptr()
        p.conv_bias = bias_.data_ptr()
        p.u = u.data_ptr()
        _lib.call("dimsum_conv_xproj_fwd", p, _stream(x))
    return result


def causal_conv1d_bwd(x, weight, bias_, dout, dx_, silu_activation):
    """causal_conv1d.cpp:349-427 -> [dx, dweight, dbias]."""
    batch, dim, seqlen, width = _checks(x, weight, bias_)
    _check(dout.dtype == x.dtype and dout.is_cuda and tuple(dout.shape) == (batch, dim, seqlen) and dout.stride(2) == 1,
           "causal_conv1d_bwd: dout must match x in dtype and shape with stride(2) == 1")
    with torch.cuda.device(x.device):
        if dx_ is not None:
            dx = dx_
            _check(dx.dtype == x.dtype and dx.is_cuda and tuple(dx.shape) == (batch, dim, seqlen) and dx.stride(2) == 1,
                   "causal_conv1d_bwd: dx must match x in dtype and shape with stride(2) == 1")
        else:
            dx = torch.empty(x.shape, device=x.device, dtype=x.dtype)
        # fp32 accumulators for the atomics, cast back at the end (causal_conv1d.cpp:402-405,425)
        dweight = torch.zeros((dim, width), device=x.device, dtype=torch.float32)
        dbias = torch.zeros((dim,), device=x.device, dtype=torch.float32) if bias_ is not None else None
        p = _lib.ConvBwdParams()
        p.batch, p.dim, p.seqlen, p.width = batch, dim, seqlen, width
        p.io_dtype, p.w_dtype, p.silu = _DT[x.dtype], _DT[weight.dtype], int(bool(silu_activation))
        p.x_batch_stride, p.x_d_stride = x.stride(0), x.stride(1)
        p.dout_batch_stride, p.dout_d_stride = dout.stride(0), dout.stride(1)
        p.dx_batch_stride, p.dx_d_stride = dx.stride(0), dx.stride(1)
        p.w_d_stride, p.w_width_stride = weight.stride(0), weight.stride(1)
        p.x, p.weight, p.dout, p.dx = x.data_ptr(), weight.data_ptr(), dout.data_ptr(), dx.data_ptr()
        p.bias = bias_.data_ptr() if bias_ is not None else None
        p.dweight = dweight.data_ptr()
        p.dbias = dbias.data_ptr() if dbias is not None else None
        _lib.call("dimsum_causal_conv1d_bwd", p, _stream(x))
    return [dx, dweight.to(weight.dtype), dbias.to(bias_.dtype) if bias_ is not None else None]


def causal_conv1d_bwd_cond(x, weight, bias_, dout, dx_, silu_activation, init_x=None):
    """causal_conv1d.cpp:429-510 -> [dx, dweight, dbias, dcond]; the conditioning input gets no gradient."""
    return causal_conv1d_bwd(x, weight, bias_, dout, dx_, silu_activation) + [None]


def causal_conv1d_update(x, conv_state, weight, bias_, silu_activation):
    raise NotImplementedError("causal_conv1d_update (autoregressive decode) is outside the DiMSUM hot path")
