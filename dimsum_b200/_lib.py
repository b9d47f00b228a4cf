"""Build and load libdimsum_b200.so (the C-ABI of include/dimsum_b200.h) through ctypes.

There is deliberately NO fallback: if the shared library is missing or a symbol is absent the import of
any op raises, so a GPU test can never pass on an eager/PyTorch path by accident.
"""
import ctypes
import os
import re
import subprocess
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
CSRC = os.path.join(_HERE, "csrc")
INCLUDE = os.path.join(_ROOT, "include")
HEADER = os.path.join(INCLUDE, "dimsum_b200.h")
LIB_PATH = os.path.join(_HERE, "libdimsum_b200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared",
]

DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into dimsum_b200/libdimsum_b200.so (in-tree)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    build_dir = os.path.join(_HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and all(os.path.getmtime(obj) > os.path.getmtime(os.path.join(CSRC, h))
                        for h in os.listdir(CSRC) if h.endswith(".cuh"))
                and os.path.getmtime(obj) > os.path.getmtime(HEADER)):
            continue
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "-Xcompiler", "-fPIC", "-I", INCLUDE, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{out}")
        if verbose and out:
            print(out)
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-Xcompiler", "-fPIC", "-o", LIB_PATH] + objs
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return LIB_PATH


# ---------------------------------------------------------------------------------------------------
# ctypes mirrors of the POD structs, generated from the header so the two can never drift apart
# ---------------------------------------------------------------------------------------------------
_CTYPE = {
    "int64_t": ctypes.c_int64,
    "float": ctypes.c_float,
}


def _parse_structs():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    structs = {}
    for body, name in re.findall(r"typedef struct \{(.*?)\}\s*(\w+);", text, flags=re.S):
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            m = re.match(r"(const\s+)?(\w+)\s+(.*)", decl)
            base, rest = m.group(2), m.group(3)
            for item in rest.split(","):
                item = item.strip()
                if item.startswith("*"):
                    fields.append((item.lstrip("* "), ctypes.c_void_p))
                else:
                    fields.append((item, _CTYPE[base]))
        structs[name] = type(name, (ctypes.Structure,), {"_fields_": fields})
    return structs


STRUCTS = _parse_structs()
ScanFwdParams = STRUCTS["dimsum_scan_fwd_params"]
ScanBwdParams = STRUCTS["dimsum_scan_bwd_params"]
ConvFwdParams = STRUCTS["dimsum_conv_fwd_params"]
ConvBwdParams = STRUCTS["dimsum_conv_bwd_params"]
ConvXprojParams = STRUCTS["dimsum_conv_xproj_params"]
AttentionParams = STRUCTS["dimsum_attention_params"]
GatherParams = STRUCTS["dimsum_gather_params"]
WaveletParams = STRUCTS["dimsum_wavelet_params"]
RowwiseParams = STRUCTS["dimsum_rowwise_params"]
RmsnormParams = STRUCTS["dimsum_rmsnorm_params"]
GeluMulParams = STRUCTS["dimsum_gelu_mul_params"]
NormModulateParams = STRUCTS["dimsum_norm_modulate_params"]
CfgEulerParams = STRUCTS["dimsum_cfg_euler_params"]
ColsumParams = STRUCTS["dimsum_colsum_params"]
GeluMulBwdParams = STRUCTS["dimsum_gelu_mul_bwd_params"]
RmsnormBwdParams = STRUCTS["dimsum_rmsnorm_bwd_params"]

ENTRY_POINTS = {
    "dimsum_selective_scan_fwd": ScanFwdParams,
    "dimsum_selective_scan_bwd": ScanBwdParams,
    "dimsum_causal_conv1d_fwd": ConvFwdParams,
    "dimsum_causal_conv1d_bwd": ConvBwdParams,
    "dimsum_conv_xproj_fwd": ConvXprojParams,
    "dimsum_attention_fwd": AttentionParams,
    "dimsum_token_gather": GatherParams,
    "dimsum_wavelet_packet_fwd": WaveletParams,
    "dimsum_wavelet_packet_inv": WaveletParams,
    "dimsum_modulate": RowwiseParams,
    "dimsum_gate_residual": RowwiseParams,
    "dimsum_add_rmsnorm": RmsnormParams,
    "dimsum_gelu_mul": GeluMulParams,
    "dimsum_norm_modulate": NormModulateParams,
    "dimsum_cfg_euler_step": CfgEulerParams,
    "dimsum_token_colsum": ColsumParams,
    "dimsum_gelu_mul_bwd": GeluMulBwdParams,
    "dimsum_add_rmsnorm_bwd": RmsnormBwdParams,
}


def declared_symbols():
    """Every function the header declares (used by the CPU test that checks the .so exports them all)."""
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(dimsum_\w+)\s*\(", text)))


_lib = None
_lock = threading.Lock()


def lib():
    """The loaded library.  Raises (never falls back) when it is missing or incomplete."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(dimsum_b200 has no CPU or PyTorch fallback)")
        handle = ctypes.CDLL(LIB_PATH)
        for name, struct in ENTRY_POINTS.items():
            fn = getattr(handle, name)
            fn.argtypes = [ctypes.POINTER(struct), ctypes.c_void_p]
            fn.restype = ctypes.c_int
        handle.dimsum_last_error.restype = ctypes.c_char_p
        handle.dimsum_last_error.argtypes = []
        handle.dimsum_abi_version.restype = ctypes.c_int
        handle.dimsum_launch_count.restype = ctypes.c_int64
        if handle.dimsum_abi_version() != 1:
            raise RuntimeError("libdimsum_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def call(name, params, stream):
    """Invoke an entry point; map status codes to the exceptions the reference raises."""
    handle = lib()
    rc = getattr(handle, name)(ctypes.byref(params), ctypes.c_void_p(stream))
    if rc == 0:
        return
    msg = handle.dimsum_last_error().decode()
    if rc == -2:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)


def launch_count():
    return int(lib().dimsum_launch_count())
