"""Import the UNMODIFIED reference Python modules from /root/reference (build container only).

TEST INFRASTRUCTURE.  Used by `oracle/make_golden.py` to produce `tests/golden/*`.
/root/reference does not exist on the GPU box, so nothing under `tests -m gpu`,
`smoke()` or `bench.py` calls this.

The two pybind extensions are replaced by empty stub modules (only the pure
PyTorch `*_ref` functions are used), `mamba_ssm/__init__.py` is bypassed because it
imports an LM wrapper needing an old `transformers`, and `pywt` is replaced by the 4
Haar taps (PyWavelets 1.6.0 `Wavelet('haar')`: dec_lo=[s,s], dec_hi=[-s,s],
rec_lo=[s,s], rec_hi=[s,-s], s=2**-0.5).
"""
import os
import sys
import types
import warnings

REFERENCE_ROOT = "/root/reference"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "dimsum"))


def load_reference():
    """Returns a namespace with the reference callables used to pin the oracle."""
    if not reference_available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    warnings.filterwarnings("ignore", category=FutureWarning)
    for name in ("selective_scan_cuda", "causal_conv1d_cuda"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "mamba_ssm" not in sys.modules:
        pkg = types.ModuleType("mamba_ssm")
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "mamba", "mamba_ssm")]
        sys.modules["mamba_ssm"] = pkg
    if "pywt" not in sys.modules:
        pw = types.ModuleType("pywt")

        class Wavelet:  # Haar only
            def __init__(self, name):
                assert name == "haar"
                s = 2.0 ** -0.5
                self.dec_lo, self.dec_hi = [s, s], [-s, s]
                self.rec_lo, self.rec_hi = [s, s], [s, -s]

        pw.Wavelet = Wavelet
        sys.modules["pywt"] = pw
    for p in (os.path.join(REFERENCE_ROOT, "causal-conv1d"), os.path.join(REFERENCE_ROOT, "dimsum")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import mamba_ssm.ops.selective_scan_interface as ssi
    from mamba_ssm.ops.selective_scan_interface import selective_scan_ref, mamba_inner_ref
    from causal_conv1d.causal_conv1d_interface import causal_conv1d_ref
    # mamba_inner_ref (:1455) calls the CUDA-backed *_fn entry points; on CPU route them to the
    # reference's own *_ref functions (module globals only -- no reference file is modified).
    ssi.causal_conv1d_fn = causal_conv1d_ref
    ssi.selective_scan_fn = selective_scan_ref
    import scanning_orders
    import wavelet_layer

    ns = types.SimpleNamespace(
        selective_scan_ref=selective_scan_ref,
        mamba_inner_ref=mamba_inner_ref,
        causal_conv1d_ref=causal_conv1d_ref,
        scanning_orders=scanning_orders,
        wavelet_layer=wavelet_layer,
    )
    return ns
