"""CPU: the reference arm of bench.py (`--impl reference`, the only arm that runs without a GPU) prints one JSON line
with the keys the driver reads, and bench.py's own arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "DiMSUM-L/2 fwd latents/s" and d["unit"] == "latents/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "latents/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None and d["data"] == "synthetic"


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour of a box without a GPU")
def test_b200_arm_fails_loudly_without_cuda():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "3", "--no-cpu-baseline"],
                         capture_output=True, text=True, cwd=ROOT, timeout=600)
    assert out.returncode != 0 and not any(l.startswith("{") for l in out.stdout.splitlines())
