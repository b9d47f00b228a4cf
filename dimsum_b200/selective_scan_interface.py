"""Operator API of the Mamba hot path, with the reference's exact names and signatures
(mamba/mamba_ssm/ops/selective_scan_interface.py): `selective_scan_fn` (:94), `mamba_inner_fn` (:1277),
`mamba_inner_fn_cond` (:1313), `mamba_inner_fn_no_out_proj[_cond]` (:1350-1452).

Everything runs on the sm_100a kernels behind the C-ABI; there is no PyTorch or CPU path here.  When autograd
is not recording (sampling) the scan skips the pre-gate `out` and checkpoint stores, which is the
"inference" byte count of SURVEY.md section 8d.
"""
import os

import torch
import torch.nn.functional as F

from . import amp, causal_conv1d_cuda, selective_scan_cuda


def fused_xproj_enabled():
    """DIMSUM_FUSED_XPROJ=0 runs conv and x_proj as two steps (the round-1 path), for A/B measurements."""
    return os.environ.get("DIMSUM_FUSED_XPROJ", "1") != "0"


def _last_contig(t):
    return t if t is None or t.stride(-1) == 1 else t.contiguous()


class SelectiveScanFn(torch.autograd.Function):
    """selective_scan_interface.py:12-91."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False,
                recording=True):
        u, delta, B, C, z = map(_last_contig, (u, delta, B, C, z))
        D = D.contiguous() if D is not None else None
        ctx.squeeze_B = B.dim() == 3
        ctx.squeeze_C = C.dim() == 3
        if ctx.squeeze_B:
            B = B.unsqueeze(1)
        if ctx.squeeze_C:
            C = C.unsqueeze(1)
        # `recording` is torch.is_grad_enabled() sampled by the caller: inside forward() grad mode is always off, and
        # ctx.needs_input_grad ignores no_grad()
        needs_grad = recording and any(ctx.needs_input_grad)
        out, x, *rest = selective_scan_cuda.fwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus,
                                                need_out=needs_grad or z is None,
                                                need_x=needs_grad or return_last_state)
        ctx.delta_softplus = delta_softplus
        ctx.has_z = z is not None
        last_state = x[:, :, -1, 1::2] if return_last_state else None  # (batch, dim, dstate)
        if not ctx.has_z:
            ctx.save_for_backward(u, delta, A, B, C, D, delta_bias, x)
            return out if not return_last_state else (out, last_state)
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, x, out)
        out_z = rest[0]
        return out_z if not return_last_state else (out_z, last_state)

    @staticmethod
    def backward(ctx, dout, *args):
        if not ctx.has_z:
            u, delta, A, B, C, D, delta_bias, x = ctx.saved_tensors
            z = out = None
        else:
            u, delta, A, B, C, D, z, delta_bias, x, out = ctx.saved_tensors
        dout = _last_contig(dout)
        du, ddelta, dA, dB, dC, dD, ddelta_bias, *rest = selective_scan_cuda.bwd(
            u, delta, A, B, C, D, z, delta_bias, dout, x, out, None, ctx.delta_softplus, False)
        dz = rest[0] if ctx.has_z else None
        dB = dB.squeeze(1) if ctx.squeeze_B else dB
        dC = dC.squeeze(1) if ctx.squeeze_C else dC
        return (du, ddelta, dA, dB, dC, dD if D is not None else None, dz,
                ddelta_bias if delta_bias is not None else None, None, None, None)


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False, return_last_state=False):
    """if return_last_state is True, returns (out, last_state); last_state has shape (batch, dim, dstate).
    The gradient of the last state is not considered in the backward pass (as in the reference)."""
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state,
                                 torch.is_grad_enabled())


def _rows_times_wt(t_bdl, weight):
    """(B, D, L) channel-major activations x (E, D) weight -> (B, L, E), i.e. F.linear(t.transpose(1, 2), weight), as one
    batched GEMM that consumes the transposed view in place (F.linear would first materialise a (B, L, D) copy)."""
    w = weight.to(t_bdl.dtype) if weight.dtype != t_bdl.dtype else weight
    return torch.bmm(t_bdl.transpose(1, 2), w.t().unsqueeze(0).expand(t_bdl.shape[0], -1, -1))


def _autocast_weights(*ws):
    """Weights in the autocast dtype: the attached bf16 shadow when there is one (amp.Bf16Shadows), else a cast."""
    if not torch.is_autocast_enabled():
        return ws
    dt = torch.get_autocast_dtype("cuda")
    out = []
    for w in ws:
        sh = amp.shadow_of(w)
        out.append(sh if sh is not None else (w.to(dt) if w is not None else None))
    return tuple(out)


class MambaInnerFn(torch.autograd.Function):
    """conv -> x_proj -> dt_proj -> scan [-> out_proj], variable B and C, checkpoint level 1.

    One class serves the reference's four composites (MambaInnerFn :579, MambaInnerFnCond :793,
    MambaInnerFnNoOutProj :174, MambaInnerFnNoOutProjCond :375): `init_states` only selects the output buffer of
    the conv and `out_proj_weight is None` skips the final projection.
    """

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias,
                A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True,
                init_states=None, has_out_proj=True, recording=True, a_arith=False):
        if B is not None or C is not None:
            raise NotImplementedError("mamba_inner_fn: only input-dependent B and C are implemented")
        if A.is_complex():
            raise NotImplementedError("mamba_inner_fn: complex A is not implemented")
        L = xz.shape[-1]
        rank = delta_proj_weight.shape[1]
        N = A.shape[-1]
        # dtypes of the master weights: backward writes their gradients in these directly from the GEMMs
        ctx.w_dtypes = tuple(w.dtype if w is not None else None for w in (x_proj_weight, delta_proj_weight, out_proj_weight))
        x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias = _autocast_weights(
            x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias)
        xz = _last_contig(xz)
        conv_w = conv1d_weight.reshape(conv1d_weight.shape[0], conv1d_weight.shape[-1])
        x, z = xz.chunk(2, dim=1)
        conv1d_bias = conv1d_bias.contiguous() if conv1d_bias is not None else None
        R, Dm = x.shape[0], x.shape[1]
        precise = x.dtype == torch.float32 and not torch.backends.cuda.matmul.allow_tf32
        if (fused_xproj_enabled() and B_proj_bias is None and C_proj_bias is None and rank % 8 == 0
                and causal_conv1d_cuda.conv_xproj_supported(x, conv_w, x_proj_weight, out=init_states, precise=precise)):
            # conv + SiLU fused in front of the x_proj contraction (tcgen05): one pass over x writes u once and emits dt
            # as the (rank, R*L) operand of the dt_proj GEMM and B / C in the scan's layout -- no second read of u, no
            # rearrange copies (selective_scan_interface.py:836-866 in one kernel)
            conv_out, dt2d, bc = causal_conv1d_cuda.conv_xproj_fwd(x, conv_w, conv1d_bias, x_proj_weight, precise,
                                                                   out=init_states, split=rank)
            Bm, Cm = bc[:, :N].unsqueeze(1), bc[:, N:].unsqueeze(1)
        else:
            conv_out = causal_conv1d_cuda.causal_conv1d_fwd_cond(x, conv_w, conv1d_bias, True, init_states)
            x_dbl = _rows_times_wt(conv_out, x_proj_weight).reshape(R * L, -1)
            dt2d = x_dbl[:, :rank].t()                                   # (rank, R*L) view
            Bm = x_dbl[:, rank:rank + N]
            Cm = x_dbl[:, rank + N:]
            if B_proj_bias is not None:
                Bm = Bm + B_proj_bias.to(Bm.dtype)
            if C_proj_bias is not None:
                Cm = Cm + C_proj_bias.to(Cm.dtype)
            Bm = Bm.view(R, L, 1, N).permute(0, 2, 3, 1).contiguous()
            Cm = Cm.view(R, L, 1, N).permute(0, 2, 3, 1).contiguous()
        # delta keeps d slowest / l fastest, the layout the scan wants (selective_scan_interface.py:837-841)
        delta = (delta_proj_weight @ dt2d).view(Dm, R, L).transpose(0, 1)
        D = D.contiguous() if D is not None else None
        needs_grad = recording and any(ctx.needs_input_grad)
        out, x_ckpt, out_z = selective_scan_cuda.fwd(conv_out, delta, A, Bm, Cm, D, z, delta_bias, delta_softplus,
                                                     need_out=needs_grad, need_x=needs_grad, a_arith=a_arith)
        ctx.delta_softplus = delta_softplus
        ctx.has_out_proj = has_out_proj
        ctx.B_bias, ctx.C_bias = B_proj_bias is not None, C_proj_bias is not None
        ctx.out_bias = out_proj_bias is not None
        if needs_grad:  # conv_out and delta are recomputed in backward (checkpoint level 1, :876-877)
            ctx.save_for_backward(xz, conv_w, conv1d_bias, dt2d, x_proj_weight, delta_proj_weight, out_proj_weight,
                                  A, Bm, Cm, D, delta_bias, x_ckpt, out)
        if not has_out_proj:
            return out_z
        y = _rows_times_wt(out_z, out_proj_weight)
        return y if out_proj_bias is None else y + out_proj_bias

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dout):
        (xz, conv_w, conv1d_bias, dt2d, x_proj_weight, delta_proj_weight, out_proj_weight, A, Bm, Cm, D, delta_bias,
         x_ckpt, out) = ctx.saved_tensors
        R, twoD, L = xz.shape
        Dm = twoD // 2
        rank = delta_proj_weight.shape[1]
        N = A.shape[-1]
        x, z = xz.chunk(2, dim=1)
        dout = _last_contig(dout)
        # conv_out is recomputed channel-major over the WHOLE batch ((Dm, R, L) memory, like delta): the scan kernels take any
        # batch / channel strides, and (Dm, R*L) views of conv_out and of the scan's du feed the x_proj GEMMs below without the
        # transposing copies a (R, Dm, L) layout would need
        conv_out = causal_conv1d_cuda.causal_conv1d_fwd_cond(
            x, conv_w, conv1d_bias, True, torch.empty((Dm, R, L), device=xz.device, dtype=xz.dtype).transpose(0, 1))
        delta = (delta_proj_weight @ dt2d).view(Dm, R, L).transpose(0, 1)
        dxz = torch.empty_like(xz)
        dx, dz = dxz.chunk(2, dim=1)
        if ctx.has_out_proj:
            dout2 = dout.reshape(R * L, -1).t()                                   # (e, R*L)
            dout_y = (out_proj_weight.t() @ dout2).view(Dm, R, L).transpose(0, 1)
        else:
            dout_y = dout
        dconv_out, ddelta, dA, dB, dC, dD, ddelta_bias, dz, out_z = selective_scan_cuda.bwd(
            conv_out, delta, A, Bm, Cm, D, z, delta_bias, dout_y, x_ckpt, out, dz, ctx.delta_softplus, True)
        dout_proj_weight = dout_proj_bias = None
        if ctx.has_out_proj:
            dout_proj_weight = amp.mm_wgrad(dout2, out_z.transpose(1, 2).reshape(R * L, Dm), ctx.w_dtypes[2])
            dout_proj_bias = dout2.sum(dim=1) if ctx.out_bias else None
        dx_dbl = torch.empty((R * L, rank + 2 * N), device=xz.device, dtype=dt2d.dtype)
        dBf = dB.squeeze(1).transpose(1, 2).reshape(R * L, N)
        dCf = dC.squeeze(1).transpose(1, 2).reshape(R * L, N)
        dx_dbl[:, rank:rank + N] = dBf
        dx_dbl[:, rank + N:] = dCf
        dB_proj_bias = dBf.sum(0) if ctx.B_bias else None
        dC_proj_bias = dCf.sum(0) if ctx.C_bias else None
        ddelta2 = ddelta.transpose(0, 1).reshape(Dm, R * L)
        ddelta_proj_weight = amp.mm_wgrad(ddelta2, dt2d.t(), ctx.w_dtypes[1])
        dx_dbl[:, :rank] = ddelta2.t() @ delta_proj_weight
        conv_t = conv_out.transpose(0, 1).reshape(Dm, R * L)                      # a view: conv_out is (Dm, R, L) memory
        dx_proj_weight = amp.mm_wgrad(dx_dbl.t(), conv_t.t(), ctx.w_dtypes[0])
        # dconv_out += x_proj_weight^T (Dm, E) @ dx_dbl^T (E, R*L): the GEMM accumulates into the scan's du in its own layout
        # (one pass; a plain `dx_dbl @ x_proj_weight` would come out token-major and need a strided add)
        xw = x_proj_weight if x_proj_weight.dtype == dx_dbl.dtype else x_proj_weight.to(dx_dbl.dtype)
        if dconv_out.dtype == dx_dbl.dtype and dconv_out.stride() == (L, R * L, 1):     # du = empty_like(u): channel-major too
            dconv_out.transpose(0, 1).view(Dm, R * L).addmm_(xw.t(), dx_dbl.t())
        else:
            dconv_out = _last_contig(dconv_out + (dx_dbl @ xw).view(R, L, Dm).transpose(1, 2))
        dx, dconv_w, dconv_b = causal_conv1d_cuda.causal_conv1d_bwd(x, conv_w, conv1d_bias, dconv_out, dx, True)
        return (dxz, dconv_w.unsqueeze(1), dconv_b if conv1d_bias is not None else None, dx_proj_weight,
                ddelta_proj_weight, dout_proj_weight, dout_proj_bias, dA, None, None, dD,
                ddelta_bias if delta_bias is not None else None, dB_proj_bias, dC_proj_bias, None, None, None, None, None)


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias,
                   A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True,
                   a_arith=False):
    """`a_arith` (extension, default off): the caller has checked that every row of A is an arithmetic progression
    (selective_scan_cuda.rows_are_arithmetic), which lets the forward scan derive its 16 decays from one exp."""
    return MambaInnerFn.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                              out_proj_bias, A, B, C, D, delta_bias, B_proj_bias, C_proj_bias, delta_softplus, None, True,
                              torch.is_grad_enabled(), a_arith)


def mamba_inner_fn_cond(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                        out_proj_bias, A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None,
                        delta_softplus=True, init_states=None):
    return MambaInnerFn.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                              out_proj_bias, A, B, C, D, delta_bias, B_proj_bias, C_proj_bias, delta_softplus,
                              init_states, True, torch.is_grad_enabled())


def mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None, C=None,
                               D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True):
    return MambaInnerFn.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, None, None, A, B, C, D,
                              delta_bias, B_proj_bias, C_proj_bias, delta_softplus, None, False, torch.is_grad_enabled())


def mamba_inner_fn_no_out_proj_cond(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None, C=None,
                                    D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True,
                                    init_states=None):
    return MambaInnerFn.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, None, None, A, B, C, D,
                              delta_bias, B_proj_bias, C_proj_bias, delta_softplus, init_states, False, torch.is_grad_enabled())
