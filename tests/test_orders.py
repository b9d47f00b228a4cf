"""CPU: the PRODUCT scan-order tables (dimsum_b200/scanning_orders.py) are bit-exact with the reference's."""
import hashlib
import json
import os

import numpy as np

from dimsum_b200 import scanning_orders as so
from golden_io import GOLDEN


def _sha(paths):
    return hashlib.sha256(np.stack(paths).astype("<i8").tobytes()).hexdigest()


def test_tables_match_golden_and_sha256():
    g = np.load(os.path.join(GOLDEN, "orders.npz"))
    sha = json.load(open(os.path.join(GOLDEN, "orders_sha256.json")))
    for name, fn in so.SCAN_ZOO.items():
        for n in (2, 4, 6, 8, 12, 16, 32, 64):
            paths = fn(n)
            assert len(paths) == 8 and all(p.dtype == np.int64 for p in paths)
            for p in paths:
                assert np.array_equal(np.sort(p), np.arange(n * n))          # valid permutation
                r = so.reverse_permut_np(p)
                assert np.array_equal(p[r], np.arange(n * n))
            assert _sha(paths) == sha[f"{name}_{n}_fwd"]
            assert _sha([so.reverse_permut_np(p) for p in paths]) == sha[f"{name}_{n}_rev"]
            if n in (4, 16, 32):
                assert np.array_equal(np.stack(paths), g[f"{name}_{n}_fwd"].astype(np.int64))


def test_window_and_implicit_orders_match_golden():
    g = np.load(os.path.join(GOLDEN, "orders.npz"))
    for grid, w in ((16, 4), (32, 8), (8, 2), (12, 3)):
        for cf in (0, 1):
            assert np.array_equal(so.window_order(grid, w, bool(cf)), g[f"window_{grid}_{w}_{cf}"].astype(np.int64))
    for grid in (16, 32):
        for tr in (0, 1):
            for rv in (0, 1):
                assert np.array_equal(so.implicit_order(grid, bool(tr), bool(rv)), g[f"implicit_{grid}_{tr}_{rv}"].astype(np.int64))
    # the four implicit spatial orders of the released config are sweep tables 0, 6, 1, 7
    sw = so.sweep_path(16)
    assert np.array_equal(so.implicit_order(16, False, False), sw[0])
    assert np.array_equal(so.implicit_order(16, False, True), sw[6])
    assert np.array_equal(so.implicit_order(16, True, False), sw[1])
    assert np.array_equal(so.implicit_order(16, True, True), sw[7])


def test_odd_grid_jpeg_rejected():
    import pytest
    with pytest.raises(ValueError):
        so.jpeg_zigzag(5)
