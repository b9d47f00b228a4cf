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


def load_reference_model_module():
    """Import the reference `models_dim` on CPU: timm 0.9.12 is absent, so the three timm classes it uses are
    restated as minimal stubs (same parameter names / maths), the Triton RMSNorm kernel is routed to the
    reference's own `rms_norm_ref`, and the Mamba slow path (use_fast_path=False) is pointed at the `*_ref` ops.
    """
    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    load_reference()
    if "models_dim" in sys.modules:
        return sys.modules["models_dim"]

    class PatchEmbed(nn.Module):  # timm.layers.PatchEmbed: Conv2d(k=stride=patch) -> (B, N, C)
        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, bias=True):
            super().__init__()
            self.patch_size = (patch_size, patch_size)
            self.num_patches = (img_size // patch_size) ** 2
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)

        def forward(self, x):
            return self.proj(x).flatten(2).transpose(1, 2)

    class Attention(nn.Module):  # timm.models.vision_transformer.Attention (fused path)
        def __init__(self, dim, num_heads=8, qkv_bias=False, **kw):
            super().__init__()
            self.num_heads, self.head_dim = num_heads, dim // num_heads
            self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
            self.proj = nn.Linear(dim, dim)

        def forward(self, x):
            B, N, C = x.shape
            q, k, v = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4).unbind(0)
            x = F.scaled_dot_product_attention(q, k, v)
            return self.proj(x.transpose(1, 2).reshape(B, N, C))

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, act_layer=nn.GELU, drop=0.0):
            super().__init__()
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.fc2 = nn.Linear(hidden_features, in_features)

        def forward(self, x):
            return self.fc2(self.act(self.fc1(x)))

    timm = types.ModuleType("timm")
    timm.__path__ = []
    for name in ("timm.models", "timm.models.vision_transformer", "timm.layers", "timm.models.layers"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    sys.modules["timm"] = timm
    vt = sys.modules["timm.models.vision_transformer"]
    vt.Attention, vt.Mlp, vt.PatchEmbed = Attention, Mlp, PatchEmbed
    sys.modules["timm.layers"].use_fused_attn = lambda: True
    for m in ("timm.layers", "timm.models.layers"):
        sys.modules[m].DropPath = nn.Identity
        sys.modules[m].trunc_normal_ = nn.init.trunc_normal_
        sys.modules[m].lecun_normal_ = nn.init.normal_
        sys.modules[m].to_2tuple = lambda v: (v, v)
    sys.path.insert(0, os.path.join(REFERENCE_ROOT, "dimsum", "pe"))

    import mamba_ssm.ops.triton.layernorm as ln
    from mamba_ssm.ops.selective_scan_interface import selective_scan_ref
    from causal_conv1d.causal_conv1d_interface import causal_conv1d_ref

    def rms_norm_fn(x, weight, bias, residual=None, prenorm=False, residual_in_fp32=False, eps=1e-6):
        return ln.rms_norm_ref(x, weight, bias, residual=residual, eps=eps, prenorm=prenorm, upcast=True)

    ln.rms_norm_fn = rms_norm_fn
    ln.RMSNorm.forward = lambda self, x, residual=None, prenorm=False, residual_in_fp32=False: rms_norm_fn(
        x, self.weight, self.bias, residual=residual, eps=self.eps, prenorm=prenorm)
    import mamba_ssm.modules.mamba_simple as ms
    ms.selective_scan_fn = selective_scan_ref
    ms.causal_conv1d_fn = causal_conv1d_ref
    import models_dim
    models_dim.rms_norm_fn = rms_norm_fn
    return models_dim
