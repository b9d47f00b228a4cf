"""Checkpoint formats of the reference, so that released weights and its training snapshots load into this repo's `DiM`
(whose module / parameter names are the reference's, dimsum_b200/models_dim.py).

* `find_model(path)`: dimsum/download.py:17-28 -- a `torch.save`d state dict, or a training checkpoint whose "ema" entry is
  preferred (then "model"); `module.` prefixes of DDP-wrapped saves are stripped.
* `save_content` / `load_content`: the resumable snapshot of dimsum/train.py:351-363 and its resume path :238-251 -- keys
  "epoch" (the NEXT epoch), "train_steps", "args", "model", "opt", "ema" -- written to `<checkpoint_dir>/content.pth`.
* `save_checkpoint`: the periodic `<epoch:07d>.pt` of train.py:366-376 (same entries without "train_steps").
* `update_ema`: train.py:55-64 as one multi-tensor lerp (ema = decay * ema + (1 - decay) * param).

Pure PyTorch host code (no kernel on this path); the training driver itself is out of scope.
"""
import os
from collections import OrderedDict

import torch


def _strip_module(sd):
    if any(k.startswith("module.") for k in sd):
        return OrderedDict((k[len("module."):] if k.startswith("module.") else k, v) for k, v in sd.items())
    return sd


def find_model(path, prefer=("ema", "model")):
    """-> state dict.  `path` is a file written by the reference (released `pytorch_model.bin`, `content.pth`, `0000100.pt`)
    or by this module.  Nothing is downloaded: a name that is not a file raises FileNotFoundError."""
    if not os.path.isfile(path):
        raise FileNotFoundError(f"find_model: no checkpoint at {path!r} (downloads are not supported)")
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    if isinstance(ckpt, dict):
        for key in prefer:
            if key in ckpt and isinstance(ckpt[key], dict):
                ckpt = ckpt[key]
                break
    if not isinstance(ckpt, dict) or not all(torch.is_tensor(v) for v in ckpt.values()):
        raise RuntimeError(f"find_model: {path!r} does not hold a state dict")
    return _strip_module(ckpt)


def load_model(model, path_or_state_dict, strict=True):
    """Load a reference checkpoint into `model`; returns `load_state_dict`'s report.  Cached scan tables / init-form flags of
    the mixers are invalidated by their `_load_from_state_dict` hook."""
    sd = find_model(path_or_state_dict) if isinstance(path_or_state_dict, (str, os.PathLike)) else _strip_module(path_or_state_dict)
    return model.load_state_dict(sd, strict=strict)


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(model.module, torch.nn.Module) else model


def save_content(checkpoint_dir, *, epoch, train_steps, args, model, opt, ema):
    """Write `<checkpoint_dir>/content.pth` like train.py:351-363 (`epoch` is the epoch just finished; epoch + 1 is stored)."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    content = {"epoch": epoch + 1, "train_steps": train_steps, "args": args, "model": _unwrap(model).state_dict(),
               "opt": opt.state_dict(), "ema": _unwrap(ema).state_dict()}
    path = os.path.join(checkpoint_dir, "content.pth")
    torch.save(content, path)
    return path


def save_checkpoint(checkpoint_dir, *, epoch, args, model, opt, ema):
    """Write `<checkpoint_dir>/<epoch:07d>.pt` like train.py:366-376."""
    os.makedirs(checkpoint_dir, exist_ok=True)
    path = os.path.join(checkpoint_dir, f"{epoch:07d}.pt")
    torch.save({"epoch": epoch + 1, "model": _unwrap(model).state_dict(), "ema": _unwrap(ema).state_dict(),
                "opt": opt.state_dict(), "args": args}, path)
    return path


def load_content(checkpoint_dir_or_file, model, opt=None, ema=None, map_location="cpu"):
    """Resume like train.py:238-251 -> (init_epoch, train_steps).  Accepts the directory holding content.pth or a file."""
    path = checkpoint_dir_or_file
    if os.path.isdir(path):
        path = os.path.join(path, "content.pth")
    ckpt = torch.load(path, map_location=map_location, weights_only=False)
    _unwrap(model).load_state_dict(_strip_module(ckpt["model"]))
    if opt is not None:
        opt.load_state_dict(ckpt["opt"])
    if ema is not None:
        _unwrap(ema).load_state_dict(_strip_module(ckpt["ema"]))
    return ckpt["epoch"], ckpt.get("train_steps", 0)


@torch.no_grad()
def update_ema(ema_model, model, decay=0.9999):
    """Step the EMA model towards the current model (train.py:55-64), all parameters in one multi-tensor call."""
    ema_params = OrderedDict(_unwrap(ema_model).named_parameters())
    src, dst = [], []
    for name, prm in _unwrap(model).named_parameters():
        dst.append(ema_params[name])
        src.append(prm.detach())
    if dst:
        torch._foreach_lerp_(dst, src, 1.0 - decay)
