"""TEST INFRASTRUCTURE -- PyTorch(CPU) restatement of the reference's pure-PyTorch oracles.

Each function restates (does not import) one reference function and cites it.  All are
ordinary differentiable torch code, so gradient parity uses autograd through them,
exactly as the reference's own tests do (mamba/tests/ops/test_selective_scan.py:97-172).
Pinned against the reference by `oracle/make_golden.py` -> `tests/golden/*.npz`.
"""
import math

import torch
import torch.nn.functional as F

from . import orders


def selective_scan_oracle(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                          return_last_state=False):
    """Sequential fp32 recurrence h_l = exp(delta_l*A) h_{l-1} + delta_l*B_l*u_l ; y_l = <C_l, h_l>.

    Restates `selective_scan_ref`, mamba/mamba_ssm/ops/selective_scan_interface.py:104-171
    (real A only; B/C either (D,N) constants, (R,N,L) or (R,G,N,L)).  Inputs are upcast to
    fp32 and the result is cast back to u.dtype once at the end, like the reference (:120-122,:170).
    """
    io_dtype = u.dtype
    uf, df = u.float(), delta.float()
    if delta_bias is not None:
        df = df + delta_bias.float().view(1, -1, 1)
    if delta_softplus:
        df = F.softplus(df)
    R, Dm, L = uf.shape
    N = A.shape[1]
    Af = A.float()

    def per_row(M):  # -> (R, Dm, N, L) view or None when constant
        if M.dim() == 2:
            return None
        M = M.float()
        if M.dim() == 3:
            return M.unsqueeze(1).expand(R, Dm, N, L)
        G = M.shape[1]
        return M.repeat_interleave(Dm // G, dim=1)

    Bv, Cv = per_row(B), per_row(C)
    h = torch.zeros(R, Dm, N, dtype=torch.float32, device=u.device)
    ys = []
    for l in range(L):
        decay = torch.exp(df[:, :, l, None] * Af[None])                      # (R, Dm, N)
        drive = df[:, :, l, None] * uf[:, :, l, None]
        drive = drive * (B.float()[None] if Bv is None else Bv[..., l])
        h = decay * h + drive
        ys.append((h * (C.float()[None] if Cv is None else Cv[..., l])).sum(-1))
    y = torch.stack(ys, dim=-1)
    if D is not None:
        y = y + uf * D.float().view(1, -1, 1)
    if z is not None:
        y = y * F.silu(z.float())
    y = y.to(io_dtype)
    return (y, h) if return_last_state else y


def causal_conv1d_oracle(x, weight, bias=None, activation=None):
    """out[b,d,l] = act(bias[d] + sum_w weight[d,w] * x[b,d,l-(W-1)+w]), zero history.

    Restates `causal_conv1d_ref`, causal-conv1d/causal_conv1d/causal_conv1d_interface.py:49-64
    (compute dtype = weight.dtype, result cast back to x.dtype).
    """
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu, or swish")
    io_dtype = x.dtype
    xw = x.to(weight.dtype)
    L = xw.shape[-1]
    W = weight.shape[1]
    xp = F.pad(xw, (W - 1, 0))
    acc = torch.zeros_like(xw)
    for w in range(W):
        acc = acc + weight[:, w].view(1, -1, 1) * xp[..., w:w + L]
    if bias is not None:
        acc = acc + bias.view(1, -1, 1)
    if activation is not None:
        acc = F.silu(acc)
    return acc.to(io_dtype)


def rms_norm_oracle(x, weight, bias=None, residual=None, eps=1e-5, prenorm=False, upcast=True):
    """Restates `rms_norm_ref`, mamba/mamba_ssm/ops/triton/layernorm.py:32-47."""
    io_dtype = x.dtype
    if upcast:
        x = x.float()
        weight = weight.float()
        residual = residual.float() if residual is not None else None
    if residual is not None:
        x = (x + residual).to(x.dtype)
    rstd = torch.rsqrt(x.square().mean(-1, keepdim=True) + eps)
    out = (x * rstd * weight + (bias if bias is not None else 0.0)).to(io_dtype)
    return (out, x) if prenorm else out


def mamba_inner_oracle(xz, conv_w, conv_b, x_proj_w, dt_proj_w, out_proj_w, out_proj_b, A, D, delta_bias,
                       perm=None, perm_rev=None):
    """conv -> x_proj -> dt_proj -> scan -> out_proj with the optional token-order gathers.

    Restates `mamba_inner_ref` (selective_scan_interface.py:1455-1561; variable B/C, softplus) wrapped in
    the gather pair of `CondMamba.forward` (mamba_simple.py:627-634,651-657).  The reference's
    conditional variant is numerically identical.
    """
    R, twoD, L = xz.shape
    Dm = twoD // 2
    N = A.shape[1]
    rank = dt_proj_w.shape[1]
    if perm is not None:
        xz = xz[:, :, perm]
    x, z = xz[:, :Dm], xz[:, Dm:]
    xc = causal_conv1d_oracle(x, conv_w.reshape(Dm, -1), conv_b, "silu")
    x_dbl = F.linear(xc.transpose(1, 2).reshape(R * L, Dm), x_proj_w)
    delta = (dt_proj_w @ x_dbl[:, :rank].t()).view(Dm, R, L).transpose(0, 1)
    Bm = x_dbl[:, rank:rank + N].view(R, L, N).transpose(1, 2).contiguous()
    Cm = x_dbl[:, rank + N:].view(R, L, N).transpose(1, 2).contiguous()
    y = selective_scan_oracle(xc, delta, A, Bm, Cm, D, z=z, delta_bias=delta_bias, delta_softplus=True)
    out = F.linear(y.transpose(1, 2), out_proj_w, out_proj_b)
    if perm_rev is not None:
        out = out[:, perm_rev, :]
    return out


# ---------------------------------------------------------------------------------------------
# Wavelet packet (2-level Haar applied to all sub-bands) with the reference's token/channel map.
# ---------------------------------------------------------------------------------------------

_HAAR_SIGNS = torch.tensor(
    [[1, 1, 1, 1],      # ll : p+q+r+s
     [1, 1, -1, -1],    # lh : rows differ
     [1, -1, 1, -1],    # hl : columns differ
     [1, -1, -1, 1]],   # hh
    dtype=torch.float32,
)


def _haar_level(img):
    """img (..., H, W) -> (..., 4, H/2, W/2), sub-bands ll, lh, hl, hh, each 0.5*(+-p +-q +-r +-s).

    wavelet_layer.py:92-114 (filters = outer products of the reversed pywt haar dec taps, applied by
    stride-2 cross-correlation, :15-18).
    """
    p = img[..., 0::2, 0::2]
    q = img[..., 0::2, 1::2]
    r = img[..., 1::2, 0::2]
    s = img[..., 1::2, 1::2]
    quad = torch.stack([p, q, r, s], dim=-3)                                 # (..., 4, h, w)
    sg = _HAAR_SIGNS.to(img.dtype).to(img.device)
    return 0.5 * torch.einsum("kj,...jhw->...khw", sg, quad)


def _haar_level_inverse(bands):
    """(..., 4, h, w) -> (..., 2h, 2w); wavelet_layer.py:36-48,68-89 (conv_transpose2d, stride 2)."""
    sg = _HAAR_SIGNS.to(bands.dtype).to(bands.device)
    quad = 0.5 * torch.einsum("kj,...khw->...jhw", sg, bands)
    *lead, _, h, w = quad.shape
    out = quad.new_zeros(*lead, 2 * h, 2 * w)
    out[..., 0::2, 0::2] = quad[..., 0, :, :]
    out[..., 0::2, 1::2] = quad[..., 1, :, :]
    out[..., 1::2, 0::2] = quad[..., 2, :, :]
    out[..., 1::2, 1::2] = quad[..., 3, :, :]
    return out


def wavelet_packet_oracle(x):
    """(R, L, C) -> (R, L, C): `WaveDiMBlock._dwt_fast`, dimsum/models_dim.py:572-586 (2 levels).

    Closed form: coef[b,c,k1,k2,h,w] = Haar_k2(Haar_k1(X[b,:,c]))/4 lands at
    token (h*4+p1)*W + w*4+p2 with p1=(c%16)//4, p2=c%4, channel (k1*4+k2)*(C/16) + c//16.
    """
    R, L, C = x.shape
    W = int(math.isqrt(L))
    assert W * W == L and W % 4 == 0 and C % 16 == 0
    img = x.transpose(1, 2).reshape(R, C, W, W)
    lvl1 = _haar_level(img)                                                  # (R, C, k1, W/2, W/2)
    lvl2 = _haar_level(lvl1) * 0.25                                          # (R, C, k1, k2, W/4, W/4)
    g = W // 4
    coef = lvl2.reshape(R, C // 16, 4, 4, 16, g, g)                          # (R, cq, p1, p2, k, h, w)
    out = coef.permute(0, 5, 2, 6, 3, 4, 1)                                  # (R, h, p1, w, p2, k, cq)
    return out.reshape(R, L, C)


def wavelet_packet_inverse_oracle(x):
    """Exact inverse: `WaveDiMBlock._idwt_fast`, dimsum/models_dim.py:588-604."""
    R, L, C = x.shape
    W = int(math.isqrt(L))
    g = W // 4
    coef = x.reshape(R, g, 4, g, 4, 16, C // 16).permute(0, 6, 2, 4, 5, 1, 3)  # (R, cq, p1, p2, k, h, w)
    lvl2 = coef.reshape(R, C, 4, 4, g, g) * 4.0
    lvl1 = _haar_level_inverse(lvl2)                                         # (R, C, k1, W/2, W/2)
    img = _haar_level_inverse(lvl1)                                          # (R, C, W, W)
    return img.reshape(R, C, L).transpose(1, 2)


def window_scan_oracle(x, w, column_first):
    """`local_scan`, dimsum/scanning_orders.py:347-367 (grid divisible by w, no flip)."""
    grid = int(math.isqrt(x.shape[1]))
    return x[:, torch.from_numpy(orders.window_order(grid, w, column_first)).to(x.device), :]


def window_unscan_oracle(x, w, column_first):
    """`local_reverse`, dimsum/scanning_orders.py:393-416."""
    grid = int(math.isqrt(x.shape[1]))
    inv = orders.invert(orders.window_order(grid, w, column_first))
    return x[:, torch.from_numpy(inv).to(x.device), :]
