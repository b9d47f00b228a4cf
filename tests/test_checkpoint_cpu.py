"""Checkpoint formats of the reference (dimsum/download.py:17-28, dimsum/train.py:55-64, 238-251, 351-376) on a small module."""
import argparse
import os

import pytest
import torch

from dimsum_b200 import checkpoint


def _net():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.SiLU(), torch.nn.Linear(16, 4))


def test_find_model_prefers_ema_and_strips_ddp_prefixes(tmp_path):
    model, ema = _net(), _net()
    with torch.no_grad():
        for p in ema.parameters():
            p.add_(1.0)
    raw = tmp_path / "pytorch_model.bin"
    torch.save(model.state_dict(), raw)
    assert all(torch.equal(v, model.state_dict()[k]) for k, v in checkpoint.find_model(str(raw)).items())
    train = tmp_path / "0000010.pt"
    torch.save({"model": {f"module.{k}": v for k, v in model.state_dict().items()}, "ema": ema.state_dict(), "epoch": 11}, train)
    sd = checkpoint.find_model(str(train))
    assert all(torch.equal(v, ema.state_dict()[k]) for k, v in sd.items())
    sd = checkpoint.find_model(str(train), prefer=("model",))
    assert set(sd) == set(model.state_dict()) and all(torch.equal(v, model.state_dict()[k]) for k, v in sd.items())
    fresh = _net()
    with torch.no_grad():
        fresh[0].weight.zero_()
    report = checkpoint.load_model(fresh, str(train))
    assert not report.missing_keys and torch.equal(fresh[0].weight, ema[0].weight)
    with pytest.raises(FileNotFoundError):
        checkpoint.find_model("DiM-L/2")
    bad = tmp_path / "bad.pt"
    torch.save({"epoch": 3}, bad)
    with pytest.raises(RuntimeError):
        checkpoint.find_model(str(bad))


def test_content_round_trip_resumes_model_optimizer_and_ema(tmp_path):
    model, ema = _net(), _net()
    opt = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=0)
    model(torch.randn(5, 8)).square().mean().backward()
    opt.step()
    checkpoint.update_ema(ema, model, decay=0.5)
    args = argparse.Namespace(model="DiM-L/2", lr=1e-4)
    path = checkpoint.save_content(str(tmp_path), epoch=6, train_steps=1234, args=args, model=model, opt=opt, ema=ema)
    assert os.path.basename(path) == "content.pth"
    raw = torch.load(path, weights_only=False)
    assert set(raw) == {"epoch", "train_steps", "args", "model", "opt", "ema"} and raw["epoch"] == 7
    model2, ema2 = _net(), _net()
    with torch.no_grad():
        for p in list(model2.parameters()) + list(ema2.parameters()):
            p.zero_()
    opt2 = torch.optim.AdamW(model2.parameters(), lr=1e-3, weight_decay=0)
    epoch, steps = checkpoint.load_content(str(tmp_path), model2, opt2, ema2)
    assert (epoch, steps) == (7, 1234)
    for a, b in zip(model.parameters(), model2.parameters()):
        assert torch.equal(a, b)
    for a, b in zip(ema.parameters(), ema2.parameters()):
        assert torch.equal(a, b)
    s1, s2 = opt.state_dict()["state"], opt2.state_dict()["state"]
    assert all(torch.equal(s1[k]["exp_avg"], s2[k]["exp_avg"]) for k in s1)
    ckpt = checkpoint.save_checkpoint(str(tmp_path), epoch=6, args=args, model=model, opt=opt, ema=ema)
    assert os.path.basename(ckpt) == "0000006.pt" and "train_steps" not in torch.load(ckpt, weights_only=False)


def test_update_ema_matches_the_reference_loop():
    model, ema, ema_ref = _net(), _net(), _net()
    with torch.no_grad():
        for p in model.parameters():
            p.add_(torch.randn_like(p))
    checkpoint.update_ema(ema, model, decay=0.9)
    with torch.no_grad():
        for (_, e), (_, p) in zip(ema_ref.named_parameters(), model.named_parameters()):
            e.mul_(0.9).add_(p.data, alpha=0.1)                  # train.py:62-64
    for a, b in zip(ema.parameters(), ema_ref.parameters()):
        assert torch.allclose(a, b, atol=1e-6)


def test_sampling_cli_bookkeeping_matches_sample_ddp(tmp_path):
    """tools/sample.py: totals, per-rank iterations and the interleaved sample index of sample_ddp.py:140-149,184, and the
    .npz assembled from the per-sample files (:36-50)."""
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import sample as cli
    assert cli.plan(50_000, 32, 8) == (50_176, 196)
    assert cli.plan(64, 32, 2) == (64, 1)
    seen = set()
    total = 0
    for _ in range(3):                                  # 3 iterations, 2 ranks, 4 per rank: every index exactly once
        for rank in range(2):
            for i in range(4):
                seen.add(cli.sample_index(i, rank, 2, total))
        total += 8
    assert seen == set(range(24))
    ns = argparse.Namespace(model="DiM-L/2", ckpt="runs/0000100.pt", cfg_scale=4.0, per_proc_batch_size=32, num_sampling_steps=250)
    assert cli.folder_name(ns) == "DiM-L-2-0000100-cfg-4.0-32-ODE-250-euler"
    d = tmp_path / "out"
    d.mkdir()
    for i in range(5):
        np.save(d / f"{i:06d}.npy", np.full((4, 2, 2), i, dtype=np.float32))
    path, shape = cli.build_npz(str(d), 4, as_images=False)
    assert shape == (4, 4, 2, 2) and np.load(path)["arr_0"][3, 0, 0, 0] == 3
