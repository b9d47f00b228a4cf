"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the DiMSUM Mamba hot path.

Nothing in the product package (`dimsum_b200/`) imports this directory.  The only
permitted users are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py`, and there only as the checker or the timed
CPU baseline -- never as a fallback for the CUDA path.

Parity status: PINNED.  Every function here was checked against the reference's own
pure-PyTorch oracles (`selective_scan_ref`, `causal_conv1d_ref`,
`mamba_inner_ref`), its `scanning_orders.py`, `wavelet_layer.py` and
`WaveDiMBlock._dwt_fast/_idwt_fast`, imported from /root/reference by
`oracle/make_golden.py`; the resulting input/output vectors are committed under
`tests/golden/` and re-checked by `tests/test_oracle_golden.py` on every run.
The reference's CUDA extensions themselves are unbuildable here (ATen/pybind
sources, sm_70..sm_90 only, no GPU in the build container), so `oracle/_ref` does
not exist; see DESIGN.md.
"""
