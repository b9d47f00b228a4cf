"""TEST INFRASTRUCTURE -- numpy restatement of the reference's token-order tables.

Closed-form restatements (not the reference's step-by-step walkers) of

* `sweep_path`    /root/reference/dimsum/scanning_orders.py:7-40
* `zigma_path`    /root/reference/dimsum/scanning_orders.py:43-78
* `jpeg_zigzag`   /root/reference/dimsum/scanning_orders.py:81-245
* `reverse_permut_np`                                      :248-253
* `local_scan` / `local_reverse` (as index tables)         :347-416
* the implicit transpose / flip orders of `DiMBlockRaw.forward`
  /root/reference/dimsum/models_dim.py:1496-1524

Pinned bit-exact against the reference by `oracle/make_golden.py`
(tests/golden/orders.npz + sha256 known answers of SURVEY.md section 8c).
Everything is int64 like the reference tables.
"""
import hashlib

import numpy as np

_CORNERS = lambda n: [(0, 0, 1, 1), (0, n - 1, 1, -1), (n - 1, 0, -1, 1), (n - 1, n - 1, -1, -1)]


def _place(n, corner, v, h):
    sr, sc, dr, dc = corner
    return (sr + dr * v) * n + sc + dc * h


def sweep_paths(n):
    """8 raster orders: 4 corners x {row-major, column-major}."""
    idx = np.arange(n)
    outer, inner = np.meshgrid(idx, idx, indexing="ij")
    outer, inner = outer.reshape(-1), inner.reshape(-1)
    paths = []
    for corner in _CORNERS(n):
        paths.append(_place(n, corner, outer, inner).astype(np.int64))  # rows outer, cols inner
        paths.append(_place(n, corner, inner, outer).astype(np.int64))  # cols outer, rows inner
    return paths


def zigma_paths(n):
    """8 boustrophedon (serpentine) orders."""
    idx = np.arange(n)
    outer, inner = np.meshgrid(idx, idx, indexing="ij")
    outer, inner = outer.reshape(-1), inner.reshape(-1)
    snake = np.where(outer % 2 == 0, inner, n - 1 - inner)
    paths = []
    for corner in _CORNERS(n):
        paths.append(_place(n, corner, outer, snake).astype(np.int64))
        paths.append(_place(n, corner, snake, outer).astype(np.int64))
    return paths


def _antidiagonal_walk(n):
    """(v, h) of the JPEG zigzag that leaves (0,0) to the right."""
    vs, hs = [], []
    for s in range(2 * n - 1):
        lo, hi = max(0, s - n + 1), min(s, n - 1)
        rng = range(hi, lo - 1, -1) if s % 2 == 0 else range(lo, hi + 1)
        for v in rng:
            vs.append(v)
            hs.append(s - v)
    return np.array(vs), np.array(hs)


def jpeg_paths(n):
    """8 JPEG zigzag orders (anti-diagonal walk; second of each pair is its transpose).

    Even n only: for odd n the reference's left-right walker stops early and returns a
    truncated, non-bijective table (scanning_orders.py:104-146); token grids are always even.
    """
    if n % 2:
        raise ValueError("jpeg order is only defined for even grids")
    v, h = _antidiagonal_walk(n)
    paths = []
    for corner in _CORNERS(n):
        paths.append(_place(n, corner, v, h).astype(np.int64))
        paths.append(_place(n, corner, h, v).astype(np.int64))
    return paths


ORDER_ZOO = {"sweep": sweep_paths, "zigma": zigma_paths, "jpeg": jpeg_paths}


def invert(perm):
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm), dtype=perm.dtype)
    return inv


def window_order(grid, w, column_first):
    """Sequence position s -> token index for `local_scan` on a grid x grid token map.

    Row-first  : s = ((hg*Wg + wg)*w + r)*w + cc
    Col-first  : s = ((wg*Hg + hg)*w + cc)*w + r
    both map to token (hg*w + r)*grid + wg*w + cc   (grid % w == 0 only).
    """
    assert grid % w == 0
    g = grid // w
    tok = np.arange(grid * grid, dtype=np.int64).reshape(g, w, g, w)  # hg r wg cc
    if column_first:
        return tok.transpose(2, 0, 3, 1).reshape(-1).copy()
    return tok.transpose(0, 2, 1, 3).reshape(-1).copy()


def implicit_spatial_order(grid, transpose, reverse):
    """Order realised by rearrange('n (h w) c -> n (w h) c') then flip(1) in DiMBlockRaw."""
    tok = np.arange(grid * grid, dtype=np.int64).reshape(grid, grid)
    seq = tok.T.reshape(-1).copy() if transpose else tok.reshape(-1).copy()
    return seq[::-1].copy() if reverse else seq


def table_sha256(paths):
    return hashlib.sha256(np.stack(paths).astype("<i8").tobytes()).hexdigest()
