"""CPU: the C-ABI shared library loads, exports every symbol the header declares, and the ctypes mirrors of the
POD structs match the header.  Validation-only calls (rejected before any launch) are exercised; no compute."""
import ctypes
import os
import re
import subprocess

import pytest

from dimsum_b200 import _lib


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    declared = _lib.declared_symbols()
    assert {"dimsum_selective_scan_fwd", "dimsum_selective_scan_bwd", "dimsum_causal_conv1d_fwd", "dimsum_causal_conv1d_bwd",
            "dimsum_token_gather", "dimsum_wavelet_packet_fwd", "dimsum_wavelet_packet_inv", "dimsum_last_error",
            "dimsum_abi_version", "dimsum_launch_count"} <= set(declared)
    for sym in declared:
        assert hasattr(handle, sym), sym
    assert _lib.lib().dimsum_abi_version() == 1


def test_library_is_sm100a_only():
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\w+", out))
    assert archs == {"sm_100a"}, archs


def test_struct_mirrors_compile_against_header(tmp_path):
    """sizeof/offsetof from gcc on the real header == ctypes layout."""
    lines = ['#include "dimsum_b200.h"', "#include <stdio.h>", "#include <stddef.h>", "int main(){"]
    for name, st in _lib.STRUCTS.items():
        lines.append(f'printf("{name} %zu\\n", sizeof({name}));')
        for fname, _ in st._fields_:
            lines.append(f'printf("{name}.{fname} %zu\\n", offsetof({name}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", _lib.INCLUDE, str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for name, st in _lib.STRUCTS.items():
        assert int(got[name]) == ctypes.sizeof(st)
        for fname, _ in st._fields_:
            assert int(got[f"{name}.{fname}"]) == getattr(st, fname).offset, (name, fname)


def test_validation_errors_map_to_reference_exceptions():
    p = _lib.ScanFwdParams()
    p.batch, p.dim, p.seqlen, p.dstate, p.n_groups = 1, 4, 8, 300, 1
    with pytest.raises(RuntimeError, match="state dimension <= 256"):
        _lib.call("dimsum_selective_scan_fwd", p, 0)
    p.dstate = 32
    with pytest.raises(NotImplementedError, match="dstate"):
        _lib.call("dimsum_selective_scan_fwd", p, 0)
    c = _lib.ConvFwdParams()
    c.batch, c.dim, c.seqlen, c.width = 1, 4, 8, 5
    with pytest.raises(RuntimeError, match="width between 2 and 4"):
        _lib.call("dimsum_causal_conv1d_fwd", c, 0)
    w = _lib.WaveletParams()
    w.batch, w.grid, w.channels = 1, 6, 32
    w.src, w.dst = 16, 32
    with pytest.raises(RuntimeError, match="multiple of 4"):
        _lib.call("dimsum_wavelet_packet_fwd", w, 0)


def test_no_cpu_fallback_in_product_package():
    """The product never imports the oracle and never branches to a CPU implementation."""
    root = os.path.dirname(_lib.__file__)
    for fn in os.listdir(root):
        if fn.endswith(".py"):
            text = open(os.path.join(root, fn)).read()
            assert "oracle" not in text.replace("# oracle", ""), fn
