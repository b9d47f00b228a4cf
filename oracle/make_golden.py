"""TEST INFRASTRUCTURE -- generate tests/golden/* by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Every vector is produced by the reference's own code (`selective_scan_ref`, `causal_conv1d_ref`,
`mamba_inner_ref`, `scanning_orders.*`, `DWT_2D/IDWT_2D` wired exactly like
`WaveDiMBlock._dwt_fast/_idwt_fast`), with inputs drawn from a seeded CPU generator.  While
generating, the restatements in `oracle/` are compared against the reference and the script
aborts on any mismatch -- that is the "pin".  bf16 tensors are stored as their raw int16 bits
with a `__bf16` name suffix (numpy has no bfloat16).
"""
import json
import math
import os
import sys

import numpy as np
import torch
from einops import rearrange

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import orders, ref_ops  # noqa: E402
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def pack(d):
    out = {}
    for k, v in d.items():
        if v is None:
            continue
        if isinstance(v, torch.Tensor):
            v = v.detach()
            if v.dtype == torch.bfloat16:
                out[k + "__bf16"] = v.contiguous().view(torch.int16).numpy()
            else:
                out[k] = v.contiguous().numpy()
        else:
            out[k] = np.asarray(v)
    return out


def relerr(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------
def gen_orders(ref):
    so = ref.scanning_orders
    zoo = {"sweep": so.sweep_path, "zigma": so.zigma_path, "jpeg": so.jpeg_zigzag}
    store, sha = {}, {}
    for n in (2, 4, 6, 8, 12, 16, 32, 64):  # odd n: the reference's jpeg walker truncates (not a permutation)
        for name, fn in zoo.items():
            ref_paths = [np.asarray(p, dtype=np.int64) for p in fn(n)]
            mine = orders.ORDER_ZOO[name](n)
            assert len(ref_paths) == len(mine) == 8
            for a, b in zip(ref_paths, mine):
                assert a.dtype == b.dtype and np.array_equal(a, b), (name, n)
            ref_rev = [so.reverse_permut_np(p) for p in ref_paths]
            for a, b in zip(ref_rev, [orders.invert(p) for p in mine]):
                assert np.array_equal(a, b), (name, n)
            sha[f"{name}_{n}_fwd"] = orders.table_sha256(ref_paths)
            sha[f"{name}_{n}_rev"] = orders.table_sha256(ref_rev)
            if n in (4, 16, 32):
                store[f"{name}_{n}_fwd"] = np.stack(ref_paths).astype(np.int16)
                store[f"{name}_{n}_rev"] = np.stack(ref_rev).astype(np.int16)
    # SURVEY.md section 8c known answers
    assert sha["sweep_16_fwd"].startswith("250d1c0a9a7fed45") and sha["jpeg_32_rev"].startswith("ac3133f44097dd89")
    # local_scan / local_reverse expressed as index tables (probe the reference with an index image)
    for grid, w in ((16, 4), (32, 8), (8, 2), (12, 3)):
        L = grid * grid
        probe = torch.arange(L, dtype=torch.float32).view(1, L, 1)
        for cf in (False, True):
            seq = so.local_scan(probe, w=w, H=grid, W=grid, column_first=cf).reshape(-1).long().numpy()
            assert np.array_equal(seq, orders.window_order(grid, w, cf)), (grid, w, cf)
            back = so.local_reverse(probe, w=w, H=grid, W=grid, column_first=cf).reshape(-1).long().numpy()
            assert np.array_equal(back, orders.invert(orders.window_order(grid, w, cf)))
            store[f"window_{grid}_{w}_{int(cf)}"] = seq.astype(np.int16)
    # implicit orders of DiMBlockRaw (models_dim.py:1496-1524): rearrange + flip
    for grid in (16, 32):
        L = grid * grid
        probe = torch.arange(L, dtype=torch.float32).view(1, L, 1)
        for tr in (False, True):
            for rv in (False, True):
                t = probe
                if tr:
                    t = rearrange(t, "n (h w) c -> n (w h) c", h=grid, w=grid)
                if rv:
                    t = t.flip(1)
                seq = t.reshape(-1).long().numpy()
                assert np.array_equal(seq, orders.implicit_spatial_order(grid, tr, rv))
                store[f"implicit_{grid}_{int(tr)}_{int(rv)}"] = seq.astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "orders.npz"), **store)
    with open(os.path.join(OUT, "orders_sha256.json"), "w") as f:
        json.dump(sha, f, indent=1, sort_keys=True)
    print("orders: ok", len(store), "tables")


# ------------------------------------------------------------------------------------------------
SCAN_CASES = [
    # name, R, D, L, N, groups(0 => 3-D B/C), dtype, has_z, has_D, has_bias, softplus, grads
    ("fp32_model_tile", 2, 16, 256, 16, 0, torch.float32, True, True, True, True, True),
    ("bf16_model_tile", 2, 16, 256, 16, 0, torch.bfloat16, True, True, True, True, True),
    ("fp32_L1024", 1, 8, 1024, 16, 0, torch.float32, True, True, True, True, False),
    ("bf16_L1024", 1, 8, 1024, 16, 0, torch.bfloat16, True, True, True, True, False),
    ("fp32_ragged", 2, 5, 37, 8, 1, torch.float32, True, True, True, True, True),
    ("fp32_noz", 2, 4, 128, 8, 1, torch.float32, False, True, True, True, True),
    ("fp32_plain", 2, 4, 64, 16, 0, torch.float32, False, False, False, False, True),
    ("fp32_groups2", 2, 8, 96, 8, 2, torch.float32, True, True, True, True, True),
    ("fp16_tile", 1, 8, 200, 16, 0, torch.float16, True, True, True, True, False),
    ("fp32_L3000", 1, 3, 3000, 4, 0, torch.float32, True, True, True, True, False),
    ("fp32_initA", 2, 8, 256, 16, 0, torch.float32, True, True, True, True, False),
    ("fp32_N1", 1, 4, 50, 1, 0, torch.float32, True, True, False, True, False),
]


def gen_scan(ref):
    g = torch.Generator().manual_seed(0)
    store = {}
    for (name, R, D, L, N, G, dt, has_z, has_D, has_bias, sp, grads) in SCAN_CASES:
        # distributions of mamba/tests/ops/test_selective_scan.py:67-95
        if name.endswith("initA"):
            A = -torch.arange(1, N + 1, dtype=torch.float32).repeat(D, 1)
        else:
            A = -0.5 * torch.rand(D, N, generator=g)
        bshape = (R, N, L) if G == 0 else (R, G, N, L)
        Bm = torch.randn(bshape, generator=g).to(dt)
        Cm = torch.randn(bshape, generator=g).to(dt)
        Dv = torch.randn(D, generator=g) if has_D else None
        z = torch.randn(R, D, L, generator=g).to(dt) if has_z else None
        bias = 0.5 * torch.rand(D, generator=g) if has_bias else None
        u = torch.randn(R, D, L, generator=g).to(dt)
        delta = (0.5 * torch.rand(R, D, L, generator=g)).to(dt)
        if name == "fp32_model_tile":          # exercise softplus over its whole range incl. the >20 branch
            delta[0, 0, :64] = torch.linspace(-30, 30, 64)
        leaves = [t for t in (u, delta, A, Bm, Cm, Dv, z, bias) if t is not None]
        if grads:
            for t in leaves:
                t.requires_grad_(True)
        out, last = ref.selective_scan_ref(u, delta, A, Bm, Cm, Dv, z=z, delta_bias=bias, delta_softplus=sp,
                                           return_last_state=True)
        mine, mine_last = ref_ops.selective_scan_oracle(u, delta, A, Bm, Cm, Dv, z=z, delta_bias=bias,
                                                        delta_softplus=sp, return_last_state=True)
        tol = 2e-6 if dt == torch.float32 else 1e-2
        assert relerr(mine, out) <= tol and relerr(mine_last, last) <= 2e-6, (name, relerr(mine, out))
        case = dict(u=u, delta=delta, A=A, B=Bm, C=Cm, D=Dv, z=z, delta_bias=bias, out=out, last_state=last,
                    softplus=int(sp))
        if grads:
            gout = torch.randn(out.shape, generator=g).to(dt)
            gr = torch.autograd.grad(out, leaves, gout, retain_graph=True)
            gm = torch.autograd.grad(mine, leaves, gout)
            names = [n for n, t in zip("u delta A B C D z delta_bias".split(), (u, delta, A, Bm, Cm, Dv, z, bias))
                     if t is not None]
            case["dout"] = gout
            for n, a, b in zip(names, gr, gm):
                assert relerr(b, a) <= (1e-5 if dt == torch.float32 else 3e-2), (name, n, relerr(b, a))
                case["d" + n] = a
        for k, v in pack(case).items():
            store[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "scan.npz"), **store)
    print("scan: ok", len(SCAN_CASES), "cases")


# ------------------------------------------------------------------------------------------------
CONV_CASES = [
    # name, R, D, L, width, dtype, wdtype, bias, silu, strided(x is half of a 2D-channel tensor)
    ("fp32_w4_silu", 2, 16, 256, 4, torch.float32, torch.float32, True, True, True),
    ("bf16_w4_silu", 2, 16, 256, 4, torch.bfloat16, torch.float32, True, True, True),
    ("bf16_w4_bf16w", 2, 8, 151, 4, torch.bfloat16, torch.bfloat16, True, True, False),
    ("fp16_w4_silu", 1, 8, 64, 4, torch.float16, torch.float32, True, True, False),
    ("fp32_w3_nobias", 2, 6, 151, 3, torch.float32, torch.float32, False, True, False),
    ("fp32_w2_linear", 2, 5, 8, 2, torch.float32, torch.float32, True, False, False),
    ("fp32_w4_L1134", 1, 3, 1134, 4, torch.float32, torch.float32, True, True, False),
    ("fp32_w4_L3", 2, 4, 3, 4, torch.float32, torch.float32, True, True, False),
]


def gen_conv(ref):
    g = torch.Generator().manual_seed(1)
    store = {}
    for (name, R, D, L, W, dt, wdt, has_bias, silu, strided) in CONV_CASES:
        if strided:
            xz = torch.randn(R, 2 * D, L, generator=g).to(dt)
            x = xz[:, :D].detach()
        else:
            x = torch.randn(R, D, L, generator=g).to(dt)
        w = torch.randn(D, W, generator=g).to(wdt)
        b = torch.randn(D, generator=g).to(wdt) if has_bias else None
        leaves = [t for t in (x, w, b) if t is not None]
        for t in leaves:
            t.requires_grad_(True)
        act = "silu" if silu else None
        out = ref.causal_conv1d_ref(x, w, b, act)
        mine = ref_ops.causal_conv1d_oracle(x, w, b, act)
        assert relerr(mine, out) <= (2e-6 if dt == torch.float32 else 1e-2), (name, relerr(mine, out))
        gout = torch.randn(out.shape, generator=g).to(dt)
        gr = torch.autograd.grad(out, leaves, gout)
        case = dict(x=x, weight=w, bias=b, out=out, dout=gout, silu=int(silu), dx=gr[0], dweight=gr[1],
                    dbias=gr[2] if has_bias else None)
        for k, v in pack(case).items():
            store[f"{name}/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "conv.npz"), **store)
    print("conv: ok", len(CONV_CASES), "cases")


# ------------------------------------------------------------------------------------------------
def ref_dwt_fast(dwt, x, levels=2):
    """Verbatim wiring of WaveDiMBlock._dwt_fast (models_dim.py:572-586) around the reference DWT_2D."""
    x = rearrange(x, "b (h w) c -> b c h w", h=int(np.sqrt(x.size(1))))
    subbands = dwt(x)
    scale = 2 ** levels
    ps = scale
    out = (dwt(subbands) / scale).chunk(ps * ps, dim=1)
    out = torch.cat([out[i % 4 * ps + i // 4] for i in range(ps * ps)], dim=1)
    return rearrange(out, "b (c p1 p2) h w -> b (h p1 w p2) c", p1=ps, p2=ps)


def ref_idwt_fast(idwt, x, levels=2):
    """Verbatim wiring of WaveDiMBlock._idwt_fast (models_dim.py:588-604)."""
    scale = 2 ** levels
    ps = scale
    lowest = int(np.sqrt(x.size(1))) // ps
    sub = rearrange(x * scale, "b (h p1 w p2) c -> b (c p1 p2) h w", p1=ps, p2=ps, h=lowest).chunk(ps * ps, dim=1)
    sub = torch.cat([sub[i % 4 * ps + i // 4] for i in range(ps * ps)], dim=1)
    out = idwt(idwt(sub))
    return rearrange(out, "b c h w -> b (h w) c")


def gen_wavelet(ref):
    g = torch.Generator().manual_seed(2)
    dwt, idwt = ref.wavelet_layer.DWT_2D("haar"), ref.wavelet_layer.IDWT_2D("haar")
    store = {}
    for name, R, grid, C in (("g16_c32", 2, 16, 32), ("g32_c16", 1, 32, 16), ("g8_c48", 2, 8, 48)):
        x = torch.randn(R, grid * grid, C, generator=g)
        y = ref_dwt_fast(dwt, x)
        mine = ref_ops.wavelet_packet_oracle(x)
        assert relerr(mine, y) <= 1e-6, (name, relerr(mine, y))
        c = torch.randn(R, grid * grid, C, generator=g)
        xi = ref_idwt_fast(idwt, c)
        mine_i = ref_ops.wavelet_packet_inverse_oracle(c)
        assert relerr(mine_i, xi) <= 1e-6, (name, relerr(mine_i, xi))
        assert relerr(ref_idwt_fast(idwt, y), x) <= 1e-5
        # integer index tables of the transform: which (token, channel) each output reads (bit-exact claim)
        for k, v in pack(dict(x=x, coef=y, c=c, recon=xi)).items():
            store[f"{name}/{k}"] = v
    # gradient of DWT == IDWT-shaped op (wavelet_layer.py:22-33): pin through autograd of the reference
    x = torch.randn(1, 64, 16, generator=g, requires_grad=True)
    y = ref_dwt_fast(dwt, x)
    gy = torch.randn(y.shape, generator=g)
    (gx,) = torch.autograd.grad(y, x, gy)
    xm = x.detach().clone().requires_grad_(True)
    (gxm,) = torch.autograd.grad(ref_ops.wavelet_packet_oracle(xm), xm, gy)
    assert relerr(gxm, gx) <= 1e-6
    for k, v in pack(dict(x=x, gy=gy, gx=gx)).items():
        store[f"grad_g8_c16/{k}"] = v
    np.savez_compressed(os.path.join(OUT, "wavelet.npz"), **store)
    print("wavelet: ok")


# ------------------------------------------------------------------------------------------------
def gen_mamba_inner(ref):
    """mamba_inner_ref (selective_scan_interface.py:1455) on a narrow mixer, with a jpeg order gather."""
    g = torch.Generator().manual_seed(3)
    R, Dm, L, N, rank, dmodel = 2, 32, 64, 16, 4, 16
    xz = torch.randn(R, 2 * Dm, L, generator=g)
    conv_w = torch.randn(Dm, 1, 4, generator=g) * 0.5
    conv_b = torch.randn(Dm, generator=g) * 0.1
    x_proj_w = torch.randn(rank + 2 * N, Dm, generator=g) / math.sqrt(Dm)
    dt_proj_w = torch.randn(Dm, rank, generator=g) / math.sqrt(rank)
    out_proj_w = torch.randn(dmodel, Dm, generator=g) / math.sqrt(Dm)
    A = -torch.exp(torch.log(torch.arange(1, N + 1, dtype=torch.float32)).repeat(Dm, 1))
    Dv = torch.ones(Dm)
    dbias = torch.rand(Dm, generator=g) - 4.0
    perm = torch.from_numpy(np.asarray(ref.scanning_orders.jpeg_zigzag(8)[3], dtype=np.int64))
    rev = torch.from_numpy(ref.scanning_orders.reverse_permut_np(perm.numpy()).astype(np.int64))
    leaves = [xz, conv_w, conv_b, x_proj_w, dt_proj_w, out_proj_w, A, Dv, dbias]
    for t in leaves:
        t.requires_grad_(True)
    xz_p = torch.gather(xz, 2, perm[None, None, :].expand_as(xz))            # mamba_simple.py:634
    out = ref.mamba_inner_ref(xz_p, conv_w, conv_b, x_proj_w, dt_proj_w, out_proj_w, None, A, None, None, Dv,
                              delta_bias=dbias, delta_softplus=True)
    out = torch.gather(out, 1, rev[None, :, None].expand_as(out))            # mamba_simple.py:657
    mine = ref_ops.mamba_inner_oracle(xz, conv_w, conv_b, x_proj_w, dt_proj_w, out_proj_w, None, A, Dv, dbias,
                                      perm=perm, perm_rev=rev)
    assert relerr(mine, out) <= 2e-6, relerr(mine, out)
    gout = torch.randn(out.shape, generator=g)
    gr = torch.autograd.grad(out, leaves, gout)
    case = dict(xz=xz, conv_w=conv_w, conv_b=conv_b, x_proj_w=x_proj_w, dt_proj_w=dt_proj_w, out_proj_w=out_proj_w,
                A=A, D=Dv, delta_bias=dbias, perm=perm, perm_rev=rev, out=out, dout=gout)
    for n, t in zip("xz conv_w conv_b x_proj_w dt_proj_w out_proj_w A D delta_bias".split(), gr):
        case["d" + n] = t
    np.savez_compressed(os.path.join(OUT, "mamba_inner.npz"), **pack(case))
    print("mamba_inner: ok")


def gen_model(ref):
    """Reference `DiM` (models_dim.py:1557) in the released wiring (block_type=combined, cond_mamba, rms_norm,
    fused_add_norm, learnable_pe, shared attention every 4 layers, scan_type=none) at toy width, run on CPU through
    the reference's slow path (`*_ref` ops).  adaLN / final layers are re-randomised -- otherwise the
    adaLN-zero init makes the output identically zero and parity vacuous."""
    from oracle.ref_loader import load_reference_model_module
    md = load_reference_model_module()
    for name, res, hidden, depth in (("toy256", 32, 64, 5), ("toy512", 64, 32, 4)):
        torch.manual_seed(11)
        m = md.DiM(img_resolution=res, in_channels=4, hidden_size=hidden, depth=depth, num_classes=10, label_dropout=0.1,
                   scan_type="none", block_type="combined", cond_mamba=True, rms_norm=True, fused_add_norm=True,
                   learnable_pe=True, use_attn_every_k_layers=4, ssm_cfg=dict(use_fast_path=False))
        m.eval()
        g = torch.Generator().manual_seed(12)
        with torch.no_grad():
            for n, prm in m.named_parameters():
                if "adaLN_modulation" in n or n.startswith("final_layer.linear"):
                    prm.copy_(torch.randn(prm.shape, generator=g) * (0.2 if "adaLN" in n else 0.05))
                if n.endswith("norm.weight") or n.endswith("norm_2.weight"):
                    prm.copy_(1.0 + 0.1 * torch.randn(prm.shape, generator=g))
        x = torch.randn(2, 4, res, res, generator=g)
        t = torch.rand(2, generator=g)
        y = torch.randint(0, 10, (2,), generator=g)
        with torch.no_grad():
            out = m(x, t, y)
            xc = torch.cat([x, x], 0)
            yc = torch.cat([y, torch.full_like(y, 10)], 0)
            out_cfg = m.forward_with_cfg(xc, torch.cat([t, t]), yc, cfg_scale=4.0)
        store = {"in/x": x.numpy(), "in/t": t.numpy(), "in/y": y.numpy(), "out/plain": out.numpy(),
                 "out/cfg4": out_cfg.numpy(), "cfg/res": np.asarray(res), "cfg/hidden": np.asarray(hidden),
                 "cfg/depth": np.asarray(depth)}
        for k, v in m.state_dict().items():
            store["sd/" + k] = v.numpy()
        np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **store)
        print("model", name, "ok; |out| max", out.abs().max().item(), "params", sum(p.numel() for p in m.parameters()))


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    ref = load_reference()
    gen_orders(ref)
    gen_scan(ref)
    gen_conv(ref)
    gen_wavelet(ref)
    gen_mamba_inner(ref)
    gen_model(ref)
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden bytes:", total)


if __name__ == "__main__":
    main()
