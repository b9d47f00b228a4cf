"""Drop-in for the reference's pybind module `selective_scan_cuda` (mamba/csrc/selective_scan/selective_scan.cpp:494-497).

`fwd` / `bwd` keep the reference's positional signatures, checks and return lists
(selective_scan.cpp:226-336, :338-492); every tensor is allocated here with torch and handed to the C-ABI
(include/dimsum_b200.h) as raw pointers + element strides on the caller's current CUDA stream.

Differences, all deliberate and visible:
  * the layout of `x` (scan_intermediates) is private to this library: (batch, dim, ceil(L/32) + 1, 2*dstate).  Record c
    holds, planar, the state after step 32c+16 in [:dstate] and after step 32c+32 in [dstate:] -- the backward's
    16-step checkpoints -- and the last record keeps the reference's interleaved convention, so
    `last_state = x[:, :, -1, 1::2]` (selective_scan_interface.py:39) still holds;
  * complex A, constant (2-D) B/C and dstate > 16 raise NotImplementedError instead of running (no fallback).
"""
import torch

from . import _lib

_DT = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}
CHUNK = 32


def _check(cond, msg):
    if not cond:
        raise RuntimeError(msg)


def _ptr(t):
    return t.data_ptr() if t is not None else None


def _stream(t):
    return torch.cuda.current_stream(t.device).cuda_stream


def _common_checks(u, delta, A, B, C, D_, z_, delta_bias_):
    _check(u.dtype in _DT, "selective_scan: input must be float32, float16 or bfloat16")
    if A.is_complex():
        raise NotImplementedError("selective_scan: complex A is not implemented in the B200 kernels")
    _check(A.dtype == torch.float32, "selective_scan: A must be float32")
    if B.dim() < 3 or C.dim() < 3:
        raise NotImplementedError("selective_scan: constant (dim, dstate) B/C is not implemented in the B200 kernels")
    _check(delta.dtype == u.dtype, "selective_scan: delta dtype must match u")
    _check(B.dtype == u.dtype and C.dtype == u.dtype, "selective_scan: variable B/C must have the input dtype")
    for name, t in (("u", u), ("delta", delta), ("A", A), ("B", B), ("C", C)):
        _check(t.is_cuda, f"selective_scan: {name} must be a CUDA tensor")
    _check(u.stride(-1) == 1 and delta.stride(-1) == 1, "selective_scan: u and delta must have stride(-1) == 1")
    batch, dim, seqlen = u.shape
    dstate = A.shape[1]
    _check(dstate <= 256, "selective_scan only supports state dimension <= 256")
    _check(tuple(delta.shape) == (batch, dim, seqlen), "selective_scan: delta shape mismatch")
    _check(tuple(A.shape) == (dim, dstate), "selective_scan: A shape mismatch")
    n_groups = B.shape[1]
    _check(B.dim() == 4 and tuple(B.shape) == (batch, n_groups, dstate, seqlen), "selective_scan: B shape mismatch")
    _check(C.dim() == 4 and tuple(C.shape) == (batch, n_groups, dstate, seqlen), "selective_scan: C shape mismatch")
    _check(B.stride(-1) == 1 and C.stride(-1) == 1, "selective_scan: B and C must have stride(-1) == 1")
    for name, t in (("D", D_), ("delta_bias", delta_bias_)):
        if t is not None:
            _check(t.dtype == torch.float32 and t.is_cuda and t.stride(-1) == 1 and tuple(t.shape) == (dim,),
                   f"selective_scan: {name} must be a contiguous float32 CUDA tensor of shape (dim,)")
    if z_ is not None:
        _check(z_.dtype == u.dtype and z_.is_cuda and z_.stride(-1) == 1 and tuple(z_.shape) == (batch, dim, seqlen),
               "selective_scan: z must match u in dtype and shape with stride(-1) == 1")
    return batch, dim, seqlen, dstate, n_groups


def rows_are_arithmetic(A, rtol=1e-6):
    """Host check (one device sync) of A[d, n] == (n + 1) * A[d, 0]: the S4D-real init form.  Callers cache the answer."""
    if A.dim() != 2 or A.shape[1] != 16:
        return False
    steps = torch.arange(1, A.shape[1] + 1, device=A.device, dtype=A.dtype)
    want = A[:, :1] * steps
    return bool(((A - want).abs() <= rtol * want.abs()).all())


def fwd(u, delta, A, B, C, D_, z_, delta_bias_, delta_softplus, *, need_out=True, need_x=True, perm=None, a_arith=False):
    """-> [out, x] (+ [out_z] when z is given).  `need_out/need_x=False` (inference) skip those stores and
    return None in their place; `perm` (int32, seqlen) folds a token-order gather of z / scatter of out_z in."""
    batch, dim, seqlen, dstate, n_groups = _common_checks(u, delta, A, B, C, D_, z_, delta_bias_)
    has_z = z_ is not None
    _check(need_out or has_z, "selective_scan: nothing to compute (no out, no z)")
    with torch.cuda.device(u.device):
        n_chunks = (seqlen + CHUNK - 1) // CHUNK + 1
        # reference: out = empty_like(delta) (inherits delta's layout), out_z = empty_like(z) (selective_scan.cpp:304,311)
        out = torch.empty_like(delta) if need_out else None
        out_z = torch.empty_like(z_) if has_z else None
        x = torch.empty((batch, dim, n_chunks, 2 * dstate), device=u.device, dtype=torch.float32) if need_x else None
        p = _lib.ScanFwdParams()
        p.batch, p.dim, p.seqlen, p.dstate, p.n_groups = batch, dim, seqlen, dstate, n_groups
        p.n_chunks, p.chunk_len = n_chunks, CHUNK
        p.io_dtype, p.delta_softplus = _DT[u.dtype], int(bool(delta_softplus))
        p.a_is_arithmetic = int(bool(a_arith))
        p.u_batch_stride, p.u_d_stride = u.stride(0), u.stride(1)
        p.delta_batch_stride, p.delta_d_stride = delta.stride(0), delta.stride(1)
        if has_z:
            p.z_batch_stride, p.z_d_stride = z_.stride(0), z_.stride(1)
            p.out_z_batch_stride, p.out_z_d_stride = out_z.stride(0), out_z.stride(1)
        if need_out:
            p.out_batch_stride, p.out_d_stride = out.stride(0), out.stride(1)
        p.A_d_stride, p.A_dstate_stride = A.stride(0), A.stride(1)
        p.B_batch_stride, p.B_group_stride, p.B_dstate_stride = B.stride(0), B.stride(1), B.stride(2)
        p.C_batch_stride, p.C_group_stride, p.C_dstate_stride = C.stride(0), C.stride(1), C.stride(2)
        p.u, p.delta, p.A, p.B, p.C = u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr()
        p.D, p.z, p.delta_bias = _ptr(D_), _ptr(z_), _ptr(delta_bias_)
        if perm is not None:
            _check(perm.dtype == torch.int32 and perm.is_cuda and perm.is_contiguous() and perm.numel() == seqlen,
                   "selective_scan: perm must be a contiguous int32 CUDA tensor of length seqlen")
        p.perm = _ptr(perm)
        p.out, p.out_z, p.x = _ptr(out), _ptr(out_z), _ptr(x)
        _lib.call("dimsum_selective_scan_fwd", p, _stream(u))
    result = [out, x]
    if has_z:
        result.append(out_z)
    return result


def bwd(u, delta, A, B, C, D_, z_, delta_bias_, dout, x_, out_, dz_, delta_softplus, recompute_out_z):
    """-> [du, ddelta, dA, dB, dC, dD, ddelta_bias] (+ [dz]) (+ [out_z])  -- selective_scan.cpp:338-492."""
    batch, dim, seqlen, dstate, n_groups = _common_checks(u, delta, A, B, C, D_, z_, delta_bias_)
    _check(dout.dtype == u.dtype and dout.is_cuda and dout.stride(-1) == 1 and tuple(dout.shape) == (batch, dim, seqlen),
           "selective_scan_bwd: dout must match u in dtype and shape with stride(-1) == 1")
    has_z = z_ is not None
    n_chunks = (seqlen + CHUNK - 1) // CHUNK + 1
    _check(x_ is not None, "selective_scan_bwd: the forward's scan_intermediates (x) are required")
    _check(x_.dtype == torch.float32 and x_.is_cuda and x_.is_contiguous()
           and tuple(x_.shape) == (batch, dim, n_chunks, 2 * dstate), "selective_scan_bwd: x has the wrong shape")
    with torch.cuda.device(u.device):
        out = dz = out_z = None
        if has_z:
            _check(out_ is not None, "selective_scan_bwd: out is required when z is given")
            out = out_
            _check(out.dtype == u.dtype and out.is_cuda and out.stride(-1) == 1 and tuple(out.shape) == (batch, dim, seqlen),
                   "selective_scan_bwd: out must match u")
            if dz_ is not None:
                dz = dz_
                _check(dz.dtype == u.dtype and dz.is_cuda and dz.stride(-1) == 1 and tuple(dz.shape) == (batch, dim, seqlen),
                       "selective_scan_bwd: dz must match u")
            else:
                dz = torch.empty_like(z_)
            if recompute_out_z:
                out_z = torch.empty_like(out)
        du = torch.empty_like(u)
        ddelta = torch.empty_like(delta)
        dA = torch.zeros_like(A)
        dB = torch.zeros(B.shape, device=B.device, dtype=torch.float32)
        dC = torch.zeros(C.shape, device=C.device, dtype=torch.float32)
        dD = torch.zeros_like(D_) if D_ is not None else None
        ddelta_bias = torch.zeros_like(delta_bias_) if delta_bias_ is not None else None
        p = _lib.ScanBwdParams()
        p.batch, p.dim, p.seqlen, p.dstate, p.n_groups = batch, dim, seqlen, dstate, n_groups
        p.n_chunks, p.chunk_len = n_chunks, CHUNK
        p.io_dtype, p.delta_softplus = _DT[u.dtype], int(bool(delta_softplus))
        p.u_batch_stride, p.u_d_stride = u.stride(0), u.stride(1)
        p.delta_batch_stride, p.delta_d_stride = delta.stride(0), delta.stride(1)
        p.dout_batch_stride, p.dout_d_stride = dout.stride(0), dout.stride(1)
        p.du_batch_stride, p.du_d_stride = du.stride(0), du.stride(1)
        p.ddelta_batch_stride, p.ddelta_d_stride = ddelta.stride(0), ddelta.stride(1)
        if has_z:
            p.z_batch_stride, p.z_d_stride = z_.stride(0), z_.stride(1)
            p.out_batch_stride, p.out_d_stride = out.stride(0), out.stride(1)
            p.dz_batch_stride, p.dz_d_stride = dz.stride(0), dz.stride(1)
            if out_z is not None:
                p.out_z_batch_stride, p.out_z_d_stride = out_z.stride(0), out_z.stride(1)
        p.A_d_stride, p.A_dstate_stride = A.stride(0), A.stride(1)
        p.B_batch_stride, p.B_group_stride, p.B_dstate_stride = B.stride(0), B.stride(1), B.stride(2)
        p.C_batch_stride, p.C_group_stride, p.C_dstate_stride = C.stride(0), C.stride(1), C.stride(2)
        p.dB_batch_stride, p.dB_group_stride, p.dB_dstate_stride = dB.stride(0), dB.stride(1), dB.stride(2)
        p.dC_batch_stride, p.dC_group_stride, p.dC_dstate_stride = dC.stride(0), dC.stride(1), dC.stride(2)
        p.u, p.delta, p.A, p.B, p.C = u.data_ptr(), delta.data_ptr(), A.data_ptr(), B.data_ptr(), C.data_ptr()
        p.D, p.z, p.delta_bias = _ptr(D_), _ptr(z_), _ptr(delta_bias_)
        p.dout, p.out, p.x = dout.data_ptr(), _ptr(out), x_.data_ptr()
        p.du, p.ddelta, p.dz, p.out_z_recompute = du.data_ptr(), ddelta.data_ptr(), _ptr(dz), _ptr(out_z)
        p.dA, p.dB, p.dC, p.dD, p.ddelta_bias = dA.data_ptr(), dB.data_ptr(), dC.data_ptr(), _ptr(dD), _ptr(ddelta_bias)
        _lib.call("dimsum_selective_scan_bwd", p, _stream(u))
    result = [du, ddelta, dA, dB.to(B.dtype), dC.to(C.dtype), dD, ddelta_bias]
    if has_z:
        result.append(dz)
    if recompute_out_z:
        result.append(out_z)
    return result
