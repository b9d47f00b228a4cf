"""CPU: the oracle restatements reproduce the vectors the unmodified reference produced (tests/golden, made by
oracle/make_golden.py).  This is the pin that lets the GPU tests trust `oracle/`."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from golden_io import GOLDEN, load, rel_err
from oracle import orders, ref_ops


def test_order_tables_bit_exact():
    g = np.load(os.path.join(GOLDEN, "orders.npz"))
    for name, fn in orders.ORDER_ZOO.items():
        for n in (4, 16, 32):
            mine = np.stack(fn(n))
            assert mine.dtype == np.int64
            assert np.array_equal(mine, g[f"{name}_{n}_fwd"].astype(np.int64))
            assert np.array_equal(np.stack([orders.invert(p) for p in fn(n)]), g[f"{name}_{n}_rev"].astype(np.int64))
    for grid, w in ((16, 4), (32, 8), (8, 2), (12, 3)):
        for cf in (0, 1):
            assert np.array_equal(orders.window_order(grid, w, bool(cf)), g[f"window_{grid}_{w}_{cf}"].astype(np.int64))
    for grid in (16, 32):
        for tr in (0, 1):
            for rv in (0, 1):
                assert np.array_equal(orders.implicit_spatial_order(grid, bool(tr), bool(rv)),
                                      g[f"implicit_{grid}_{tr}_{rv}"].astype(np.int64))


def test_order_sha256_known_answers():
    sha = json.load(open(os.path.join(GOLDEN, "orders_sha256.json")))
    # SURVEY.md section 8c (first 16 hex digits), independent of this repo's generator
    known = {"sweep_16_fwd": "250d1c0a9a7fed45", "sweep_16_rev": "ac9a8d4830f07da1", "zigma_16_fwd": "dfe51cdf56197994",
             "zigma_16_rev": "523c5bd8c422e8d6", "jpeg_16_fwd": "a7b5aa963198bac1", "jpeg_16_rev": "535aaf883550bfb3",
             "sweep_32_fwd": "85da532cebf59ba7", "sweep_32_rev": "ba59374f3abbe06a", "zigma_32_fwd": "01b6ef874ac9cd89",
             "zigma_32_rev": "66e222b37c7d6577", "jpeg_32_fwd": "9427ae6f06d7c687", "jpeg_32_rev": "ac3133f44097dd89"}
    for k, v in known.items():
        assert sha[k].startswith(v), k
    for key, digest in sha.items():
        name, n, direction = key.split("_")
        paths = orders.ORDER_ZOO[name](int(n))
        if direction == "rev":
            paths = [orders.invert(p) for p in paths]
        assert orders.table_sha256(paths) == digest, key
    assert orders.jpeg_paths(4)[0].tolist() == [0, 1, 4, 8, 5, 2, 3, 6, 9, 12, 13, 10, 7, 11, 14, 15]


@pytest.mark.parametrize("case", sorted(load("scan.npz")))
def test_scan_oracle_matches_reference(case):
    c = load("scan.npz")[case]
    out, last = ref_ops.selective_scan_oracle(c["u"], c["delta"], c["A"], c["B"], c["C"], c.get("D"), z=c.get("z"),
                                              delta_bias=c.get("delta_bias"), delta_softplus=bool(c["softplus"]),
                                              return_last_state=True)
    t = 2e-6 if c["u"].dtype == torch.float32 else 1e-2
    assert rel_err(out, c["out"]) <= t
    assert rel_err(last, c["last_state"]) <= 2e-6


@pytest.mark.parametrize("case", ["fp32_model_tile", "fp32_ragged", "fp32_groups2"])
def test_scan_oracle_grads_match_reference(case):
    c = load("scan.npz")[case]
    names = [n for n in "u delta A B C D z delta_bias".split() if n in c]
    leaves = {n: c[n].clone().requires_grad_(True) for n in names}
    out = ref_ops.selective_scan_oracle(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves.get("D"),
                                        z=leaves.get("z"), delta_bias=leaves.get("delta_bias"),
                                        delta_softplus=bool(c["softplus"]))
    grads = torch.autograd.grad(out, [leaves[n] for n in names], c["dout"])
    for n, g in zip(names, grads):
        assert rel_err(g, c["d" + n]) <= 1e-5, n


@pytest.mark.parametrize("case", sorted(load("conv.npz")))
def test_conv_oracle_matches_reference(case):
    c = load("conv.npz")[case]
    out = ref_ops.causal_conv1d_oracle(c["x"], c["weight"], c.get("bias"), "silu" if c["silu"] else None)
    assert rel_err(out, c["out"]) <= (2e-6 if c["x"].dtype == torch.float32 else 1e-2)


def test_wavelet_oracle_matches_reference():
    g = load("wavelet.npz")
    for case in ("g16_c32", "g32_c16", "g8_c48"):
        c = g[case]
        assert rel_err(ref_ops.wavelet_packet_oracle(c["x"]), c["coef"]) <= 1e-6
        assert rel_err(ref_ops.wavelet_packet_inverse_oracle(c["c"]), c["recon"]) <= 1e-6
        assert rel_err(ref_ops.wavelet_packet_inverse_oracle(ref_ops.wavelet_packet_oracle(c["x"])), c["x"]) <= 1e-5


def test_mamba_inner_oracle_matches_reference():
    c = load("mamba_inner.npz")
    out = ref_ops.mamba_inner_oracle(c["xz"], c["conv_w"], c["conv_b"], c["x_proj_w"], c["dt_proj_w"], c["out_proj_w"],
                                     None, c["A"], c["D"], c["delta_bias"], perm=c["perm"], perm_rev=c["perm_rev"])
    assert rel_err(out, c["out"]) <= 2e-6


@pytest.mark.parametrize("name", ["toy256", "toy512"])
def test_model_oracle_matches_reference(name):
    from oracle import ref_model
    raw = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    sd = {k[3:]: torch.from_numpy(raw[k].copy()) for k in raw.files if k.startswith("sd/")}
    x, t, y = (torch.from_numpy(raw[f"in/{k}"]) for k in ("x", "t", "y"))
    with torch.no_grad():
        out = ref_model.dim_forward_oracle(sd, x, t, y)
        cfg = ref_model.dim_forward_with_cfg_oracle(sd, torch.cat([x, x]), torch.cat([t, t]),
                                                    torch.cat([y, torch.full_like(y, 10)]), 4.0)
    assert rel_err(out, torch.from_numpy(raw["out/plain"])) <= 2e-6
    assert rel_err(cfg, torch.from_numpy(raw["out/cfg4"])) <= 2e-6
