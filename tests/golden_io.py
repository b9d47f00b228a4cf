"""Load tests/golden/*.npz (bf16 tensors are stored as int16 bit patterns with a `__bf16` suffix)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    """-> {case: {tensor_name: torch.Tensor | python scalar}} for case-structured files, or a flat dict."""
    raw = np.load(os.path.join(GOLDEN, name))
    out = {}
    for key in raw.files:
        arr = raw[key]
        case, _, leaf = key.rpartition("/")
        if leaf.endswith("__bf16"):
            leaf = leaf[: -len("__bf16")]
            val = torch.from_numpy(arr.copy()).view(torch.bfloat16)
        elif arr.ndim == 0:
            val = arr.item()
        else:
            val = torch.from_numpy(arr.copy())
        if case:
            out.setdefault(case, {})[leaf] = val
        else:
            out[leaf] = val
    return out


def rel_err(a, b):
    """max|a-b| / max|b| after upcast -- the parity metric of SURVEY.md section 8d."""
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# north-star tolerances: fp32 1e-5, 16-bit 2e-2
def tol(dtype):
    return 1e-5 if dtype == torch.float32 else 2e-2
