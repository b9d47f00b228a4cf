"""DiM-L/2-sized parity with the SAME-DEVICE oracle of SURVEY.md section 8c.

The oracle (`oracle.ref_model`: reference ops restated as plain torch ops, pinned against the unmodified reference by
tests/golden) is fed CUDA tensors, TF32 off, math SDPA, so dense layers run the same cuBLAS fp32 GEMMs on both sides and only
this repo's kernels (scan, conv, wavelet, order folding, norm / modulate / gate glue) differ:

  * one DiM-L/2 (459.9 M parameters, random init, adaLN layers re-randomised) forward and a CFG forward;
  * the north star's fixed-seed criterion: CFG Euler sampling on the 250-point grid, final latent vs the oracle sampler.
    The ACHIEVED relative error is printed and recorded in profiles/; the assert uses the measured bound with the
    explanation below rather than silently loosening a toy test;
  * bf16-autocast forward + backward of the toy reference model against the oracle under the same autocast (<= 2e-2).
"""
import json
import os

import numpy as np
import pytest
import torch

from golden_io import GOLDEN, rel_err

pytestmark = pytest.mark.gpu

STEPS = int(os.environ.get("DIMSUM_SAMPLING_PARITY_STEPS", "250"))


@pytest.fixture()
def strict_fp32():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


def _math_sdpa():
    from torch.nn.attention import SDPBackend, sdpa_kernel
    return sdpa_kernel(SDPBackend.MATH)


@pytest.fixture(scope="module")
def dim_l2():
    from dimsum_b200.models_dim import DiM_models
    torch.manual_seed(0)
    with torch.device("cuda"):
        model = DiM_models["DiM-L/2"](img_resolution=32, in_channels=4, num_classes=1000, label_dropout=0.1)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if "adaLN_modulation" in n or n.startswith("final_layer.linear"):
                p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(p.device))
            if n.endswith("A_log"):      # a trained checkpoint does not keep the S4D-real init form: perturb it (general-A kernel)
                p.add_((torch.rand(p.shape, generator=g) * 0.4 - 0.2).to(p.device))
    model = model.cuda().eval()
    sd = {k: v.detach() for k, v in model.state_dict().items()}
    return model, sd


def test_dim_l2_forward_matches_same_device_oracle(dim_l2, strict_fp32):
    from oracle import ref_model
    model, sd = dim_l2
    assert sum(p.numel() for p in model.parameters()) > 4.5e8
    g = torch.Generator().manual_seed(2)
    x = torch.randn(2, 4, 32, 32, generator=g).cuda()
    t = torch.rand(2, generator=g).cuda()
    y = torch.randint(0, 1000, (2,), generator=g).cuda()
    with torch.no_grad():
        got = model(x, t, y)
        got_cfg = model.forward_with_cfg(torch.cat([x, x]), torch.cat([t, t]), torch.cat([y, torch.full_like(y, 1000)]), cfg_scale=4.0)
        with _math_sdpa():
            want = ref_model.dim_forward_oracle(sd, x, t, y)
            want_cfg = ref_model.dim_forward_with_cfg_oracle(sd, torch.cat([x, x]), torch.cat([t, t]),
                                                             torch.cat([y, torch.full_like(y, 1000)]), 4.0)
    e, e_cfg = rel_err(got, want), rel_err(got_cfg, want_cfg)
    print(f"DiM-L/2 forward rel err vs same-device oracle: {e:.3e} (cfg {e_cfg:.3e})")
    assert e <= 1e-5, e
    assert e_cfg <= 1e-5, e_cfg


def test_dim_l2_fixed_seed_cfg_sampling_final_latent(dim_l2, strict_fp32):
    """North star: 'the same final latent within that tolerance after a fixed-seed sampling run' -- DiM-L/2, CFG 4.0, fixed-grid
    Euler on linspace(0, 1, 250) (249 evaluations), one latent (2 CFG rows), this repo's sampler vs the same-device oracle
    sampler from the same noise, label and weights.  Per evaluation the two forwards agree to ~1e-6; the ODE integrates
    those differences over 249 steps, so the final-latent error is reported as measured."""
    from dimsum_b200.sampler import sample_cfg
    from oracle import ref_model
    model, sd = dim_l2
    g = torch.Generator().manual_seed(123)
    z = torch.randn(1, 4, 32, 32, generator=g).cuda()
    y = torch.randint(0, 1000, (1,), generator=g).cuda()
    got = sample_cfg(model, z, y, cfg_scale=4.0, num_steps=STEPS)
    with torch.no_grad(), _math_sdpa():
        want = ref_model.euler_sample_oracle(sd, z, y, 4.0, num_steps=STEPS, null_class=1000)
    e = rel_err(got, want)
    l2 = ((got - want).norm() / want.norm()).item()
    print(f"DiM-L/2 fixed-seed CFG Euler ({STEPS}-point grid) final latent: max-norm rel err {e:.3e}, l2 rel err {l2:.3e}")
    out_dir = os.path.join(os.path.dirname(GOLDEN), "..", "gpurun_out")
    try:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, "sampling_parity.json"), "w") as f:
            json.dump({"model": "DiM-L/2", "grid_points": STEPS, "cfg_scale": 4.0, "latents": 1,
                       "rel_err_maxnorm": e, "rel_err_l2": l2, "oracle": "same-device (CUDA tensor ops, TF32 off, math SDPA)"}, f)
    except OSError:
        pass
    assert torch.isfinite(got).all()
    assert e <= 1e-5, e


def _toy(name="toy256"):
    raw = np.load(os.path.join(GOLDEN, f"model_{name}.npz"))
    sd = {k[3:]: torch.from_numpy(raw[k].copy()) for k in raw.files if k.startswith("sd/")}
    return raw, sd


@pytest.mark.parametrize("mode", ["bf16", "bf16_shadows", "fp32"])
def test_recorded_forward_and_backward_match_oracle(mode):
    """configs[4] is a bf16-autocast config: the recorded (training) forward and ALL parameter gradients of the toy reference
    model vs autograd through the oracle on the same device -- under bf16 autocast (plain, and with the bf16 weight shadows of
    dimsum_b200/amp.py) at the bf16 tolerance, and in strict fp32 at 1e-4 (orders folded into the glue kernels' backward)."""
    from dimsum_b200 import amp
    from dimsum_b200.models_dim import DiM
    from oracle import ref_model
    raw, sd = _toy()
    m = DiM(img_resolution=int(raw["cfg/res"]), in_channels=4, hidden_size=int(raw["cfg/hidden"]), depth=int(raw["cfg/depth"]),
            num_classes=10, label_dropout=0.1, use_attn_every_k_layers=4)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().train()
    x, t, y = (torch.from_numpy(raw[f"in/{k}"]).cuda() for k in ("x", "t", "y"))
    g = torch.Generator().manual_seed(5)
    dout = torch.randn(raw["out/plain"].shape, generator=g).cuda()
    m.y_embedder.dropout_prob = 0.0                                           # deterministic labels
    low = mode != "fp32"
    shadows = amp.Bf16Shadows(m) if mode == "bf16_shadows" else None
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = low
    try:
        ctx = lambda: torch.autocast("cuda", dtype=torch.bfloat16, enabled=low)
        with ctx():
            out = m(x, t, y)
        (out.float() * dout).sum().backward()
        ours = {n: p.grad for n, p in m.named_parameters() if p.grad is not None}
        assert all(gr.dtype == torch.float32 for gr in ours.values())

        leaves = {k: v.cuda().clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        with ctx():
            want = ref_model.dim_forward_oracle(leaves, x, t, y)
        (want.float() * dout).sum().backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
        if shadows is not None:
            shadows.detach()
    tol = 2e-2 if low else 1e-4
    e = rel_err(out, want)
    print(f"{mode} recorded forward rel err vs oracle: {e:.3e}")
    assert e <= (tol if low else 5e-5), e
    worst = ("", 0.0)
    num = den = 0.0
    for n, gr in ours.items():
        ref = leaves[n].grad
        if ref is None:
            continue
        num += (gr.float() - ref.float()).square().sum().item()
        den += ref.float().square().sum().item()
        err = rel_err(gr, ref)
        if err > worst[1]:
            worst = (n, err)
    total = (num / den) ** 0.5
    print(f"{mode} gradients: global l2 rel err {total:.3e}; worst per-parameter max-norm rel err {worst[1]:.3e} ({worst[0]})")
    assert total <= tol, total
    # every trainable parameter that the oracle gives a gradient also got one here
    assert all(n in ours for n, v in leaves.items() if v.grad is not None and "cond_proj" not in n)


@pytest.mark.parametrize("shadows", [False, True])
def test_bf16_sampling_forward_matches_oracle_under_autocast(shadows):
    """The `--dtype bf16` sampling path: eval forward and forward_with_cfg under no_grad + bf16 autocast (glue kernels emitting
    the GEMM dtype, SiLU(c) cast once, optionally the bf16 weight shadows) against the oracle under the same autocast."""
    from dimsum_b200 import amp
    from dimsum_b200.models_dim import DiM
    from oracle import ref_model
    raw, sd = _toy()
    m = DiM(img_resolution=int(raw["cfg/res"]), in_channels=4, hidden_size=int(raw["cfg/hidden"]), depth=int(raw["cfg/depth"]),
            num_classes=10, label_dropout=0.1, use_attn_every_k_layers=4)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval()
    sh = amp.Bf16Shadows(m) if shadows else None
    x, t, y = (torch.from_numpy(raw[f"in/{k}"]).cuda() for k in ("x", "t", "y"))
    leaves = {k: v.cuda() for k, v in sd.items()}
    try:
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            out = m(x, t, y)
            cfg = m.forward_with_cfg(torch.cat([x, x]), torch.cat([t, t]), torch.cat([y, torch.full_like(y, 10)]), cfg_scale=4.0)
            want = ref_model.dim_forward_oracle(leaves, x, t, y)
    finally:
        if sh is not None:
            sh.detach()
    e = rel_err(out, want)
    print(f"bf16 sampling forward (shadows={shadows}) rel err vs oracle under autocast: {e:.3e}")
    assert e <= 2e-2, e
    assert rel_err(cfg, torch.from_numpy(raw["out/cfg4"]).cuda()) <= 5e-2
