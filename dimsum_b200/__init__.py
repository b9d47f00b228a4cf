"""dimsum_b200 -- B200-native (sm_100a) implementation of the DiMSUM Mamba hot path.

Public operator API (same names and signatures as the reference):
    selective_scan_fn, mamba_inner_fn, mamba_inner_fn_cond, mamba_inner_fn_no_out_proj[_cond]
    causal_conv1d_fn
    scanning_orders (sweep_path, zigma_path, jpeg_zigzag, reverse_permut_np, local_scan, local_reverse)
    wavelet_packet, wavelet_packet_inverse
All of them launch hand-written CUDA kernels through the C-ABI in include/dimsum_b200.h; there is no CPU,
PyTorch-eager or Triton fallback -- importing an op without the built library raises.
"""
from .causal_conv1d_interface import causal_conv1d_fn  # noqa: F401
from .selective_scan_interface import (  # noqa: F401
    mamba_inner_fn,
    mamba_inner_fn_cond,
    mamba_inner_fn_no_out_proj,
    mamba_inner_fn_no_out_proj_cond,
    selective_scan_fn,
)
from .wavelet import wavelet_packet, wavelet_packet_inverse  # noqa: F401
from . import scanning_orders  # noqa: F401

__version__ = "0.1.0"
