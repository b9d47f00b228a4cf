"""Token scan orders of DiMSUM (reference: dimsum/scanning_orders.py), as permutation tables + gather kernels.

Same public names as the reference: `sweep_path` (:7), `zigma_path` (:43), `jpeg_zigzag` (:81),
`reverse_permut_np` (:248), `local_scan` (:347), `local_reverse` (:393), `SCAN_ZOO` (:419).  The tables are
int64 numpy arrays, bit-identical to the reference's (tests/test_orders.py against tests/golden/orders.npz and the
sha256 known answers).  The data movement itself never materialises through PyTorch indexing: it is either a
`token_gather` launch (16-byte coalesced row copies) or, inside the model, folded into a neighbouring kernel
(`perm` argument of the conv / scan, `pos` argument of the wavelet kernels).
"""
import functools
import math

import numpy as np
import torch

from . import _lib

_DT = {torch.float32: _lib.DTYPE_F32, torch.float16: _lib.DTYPE_F16, torch.bfloat16: _lib.DTYPE_BF16}


def _eight_paths(N, cell_sequence):
    """cell_sequence(N, swap) -> list of (v, h) steps from the (0,0) corner; mirrored into the 4 corners."""
    paths = []
    for start_row, start_col, dir_row, dir_col in ((0, 0, 1, 1), (0, N - 1, 1, -1), (N - 1, 0, -1, 1), (N - 1, N - 1, -1, -1)):
        for swap in (False, True):
            paths.append(np.array([(start_row + dir_row * v) * N + start_col + dir_col * h
                                   for v, h in cell_sequence(N, swap)], dtype=np.int64))
    return paths


def _raster(N, swap):
    return [((i, j) if not swap else (j, i)) for i in range(N) for j in range(N)]


def _serpentine(N, swap):
    cells = []
    for i in range(N):
        for j in range(N):
            k = j if i % 2 == 0 else N - 1 - j
            cells.append((i, k) if not swap else (k, i))
    return cells


def _diagonal(N, swap):
    if N % 2:
        raise ValueError("jpeg_zigzag: the reference order is only a permutation for even grids")
    cells = []
    for s in range(2 * N - 1):
        span = range(max(0, s - N + 1), min(s, N - 1) + 1)
        for v in (reversed(span) if s % 2 == 0 else span):
            cells.append((v, s - v) if not swap else (s - v, v))
    return cells


def sweep_path(N):
    """Mamba's sweep scan: 4 corners x {row-major, column-major}."""
    return _eight_paths(N, _raster)


def zigma_path(N):
    """Zigma's continuity (boustrophedon) scan."""
    return _eight_paths(N, _serpentine)


def jpeg_zigzag(N):
    """JPEG zigzag over anti-diagonals."""
    return _eight_paths(N, _diagonal)


SCAN_ZOO = {"sweep": sweep_path, "zigma": zigma_path, "jpeg": jpeg_zigzag}


def reverse_permut_np(permutation):
    permutation = np.asarray(permutation)
    reverse = np.zeros(len(permutation), dtype=np.int64)
    reverse[permutation] = np.arange(len(permutation), dtype=np.int64)
    return reverse


def window_order(grid, w, column_first=False):
    """Sequence position -> token index of `local_scan(w, H=W=grid)` (grid % w == 0)."""
    if grid % w:
        raise NotImplementedError("window scan with padding (grid % w != 0) is not used by DiMSUM")
    g = grid // w
    order = np.empty(grid * grid, dtype=np.int64)
    s = 0
    if not column_first:
        for hg in range(g):
            for wg in range(g):
                for r in range(w):
                    for c in range(w):
                        order[s] = (hg * w + r) * grid + wg * w + c
                        s += 1
    else:
        for wg in range(g):
            for hg in range(g):
                for c in range(w):
                    for r in range(w):
                        order[s] = (hg * w + r) * grid + wg * w + c
                        s += 1
    return order


def implicit_order(grid, transpose=False, reverse=False):
    """Order realised by DiMBlockRaw's rearrange('n (h w) c -> n (w h) c') + flip(1) (models_dim.py:1498-1507)."""
    tok = np.arange(grid * grid, dtype=np.int64).reshape(grid, grid)
    seq = (tok.T if transpose else tok).reshape(-1)
    return seq[::-1].copy() if reverse else seq.copy()


@functools.lru_cache(maxsize=256)
def _device_table(key, device):
    kind = key[0]
    if kind == "window":
        table = window_order(key[1], key[2], key[3])
    elif kind == "window_inv":
        table = reverse_permut_np(window_order(key[1], key[2], key[3]))
    else:
        raise KeyError(kind)
    return torch.from_numpy(table.astype(np.int32)).to(device)


def as_index(table, device):
    """int64 numpy / torch table -> contiguous int32 CUDA tensor for the kernels."""
    if isinstance(table, np.ndarray):
        table = torch.from_numpy(table)
    return table.to(device=device, dtype=torch.int32).contiguous()


def token_gather(x, index):
    """out[b, l, :] = x[b, index[l], :] for token-major (batch, seqlen, channels) tensors -- one coalesced kernel."""
    if x.dim() != 3 or x.stride(2) != 1:
        raise RuntimeError("token_gather: x must be (batch, seqlen, channels) with channel stride 1")
    if index.dtype != torch.int32 or not index.is_cuda or index.numel() != x.shape[1]:
        raise RuntimeError("token_gather: index must be an int32 CUDA tensor of length seqlen")
    out = torch.empty(x.shape, device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        p = _lib.GatherParams()
        p.batch, p.seqlen, p.channels, p.dtype = x.shape[0], x.shape[1], x.shape[2], _DT[x.dtype]
        p.src_batch_stride, p.src_token_stride = x.stride(0), x.stride(1)
        p.dst_batch_stride, p.dst_token_stride = out.stride(0), out.stride(1)
        p.src, p.index, p.dst = x.data_ptr(), index.data_ptr(), out.data_ptr()
        _lib.call("dimsum_token_gather", p, torch.cuda.current_stream(x.device).cuda_stream)
    return out


class _TokenGatherFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, index, inverse):
        ctx.save_for_backward(inverse)
        return token_gather(x, index)

    @staticmethod
    def backward(ctx, g):
        (inverse,) = ctx.saved_tensors
        return token_gather(g.contiguous(), inverse), None, None


def permute_tokens(x, index, inverse):
    """Differentiable token gather; `inverse` is the inverse permutation (its gather is the gradient)."""
    return _TokenGatherFn.apply(x, index, inverse)


def local_scan(x, w=7, H=14, W=14, flip=False, column_first=False):
    """Local windowed scan (LocalMamba), x: [B, L, C] -> [B, L, C]; reference scanning_orders.py:347-367."""
    if H != W or flip:
        raise NotImplementedError("local_scan: only square grids without flip are used by DiMSUM")
    fwd = _device_table(("window", H, w, bool(column_first)), x.device)
    inv = _device_table(("window_inv", H, w, bool(column_first)), x.device)
    return permute_tokens(x.contiguous(), fwd, inv)


def local_reverse(x, w=7, H=14, W=14, flip=False, column_first=False):
    """Inverse of `local_scan`; reference scanning_orders.py:393-416."""
    if H != W or flip:
        raise NotImplementedError("local_reverse: only square grids without flip are used by DiMSUM")
    fwd = _device_table(("window", H, w, bool(column_first)), x.device)
    inv = _device_table(("window_inv", H, w, bool(column_first)), x.device)
    return permute_tokens(x.contiguous(), inv, fwd)
