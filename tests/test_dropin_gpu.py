"""The pybind-surface shims driven with EXACTLY the positional argument lists of the reference's call sites
(mamba/mamba_ssm/ops/selective_scan_interface.py:836 `causal_conv1d_fwd_cond(x, w, b, True, init_states)`, :872
`fwd(conv1d_out, delta, A, B, C, D, z, delta_bias, delta_softplus)`, :929 `causal_conv1d_fwd(x, w, b, True)`, :938
`bwd(conv1d_out, delta, A, B, C, D, z, delta_bias, dout_y, scan_intermediates, out, dz, delta_softplus, True)`, :985
`causal_conv1d_bwd(x, w, b, dconv1d_out, dx, True)`), i.e. a re-enactment of `MambaInnerFnCond.forward/backward` on top of
`dimsum_b200.selective_scan_cuda` / `dimsum_b200.causal_conv1d_cuda` used as drop-ins for the pybind modules.  Results are
checked against the reference's own `mamba_inner_ref` golden (tests/golden/mamba_inner.npz, made from the unmodified
reference by oracle/make_golden.py)."""
import pytest
import torch
import torch.nn.functional as F

from golden_io import load, rel_err

pytestmark = pytest.mark.gpu


def test_reference_call_sequence_on_the_shims_matches_mamba_inner_ref_golden():
    import dimsum_b200.causal_conv1d_cuda as causal_conv1d_cuda        # installed as sys.modules["causal_conv1d_cuda"]
    import dimsum_b200.selective_scan_cuda as selective_scan_cuda      # installed as sys.modules["selective_scan_cuda"]
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        c = load("mamba_inner.npz")
        t = {n: c[n].cuda() for n in "xz conv_w conv_b x_proj_w dt_proj_w out_proj_w A D delta_bias".split()}
        perm, rev = c["perm"].cuda(), c["perm_rev"].cuda()
        xz = torch.gather(t["xz"], 2, perm[None, None, :].expand_as(t["xz"])).contiguous()
        R, twoD, L = xz.shape
        delta_rank, d_state = t["dt_proj_w"].shape[1], t["A"].shape[1]
        # ---- forward, selective_scan_interface.py:833-899
        conv1d_weight = t["conv_w"].reshape(t["conv_w"].shape[0], -1)
        x, z = xz.chunk(2, dim=1)
        init_states = torch.full_like(x, 7.0)                               # the "cond" buffer: overwritten (Q1)
        conv1d_out = causal_conv1d_cuda.causal_conv1d_fwd_cond(x, conv1d_weight, t["conv_b"], True, init_states)
        assert conv1d_out.data_ptr() == init_states.data_ptr()
        x_dbl = F.linear(conv1d_out.transpose(1, 2).reshape(R * L, -1), t["x_proj_w"])
        delta = (t["dt_proj_w"] @ x_dbl[:, :delta_rank].t()).view(-1, R, L).transpose(0, 1)
        B = x_dbl[:, delta_rank:delta_rank + d_state].reshape(R, L, 1, d_state).permute(0, 2, 3, 1).contiguous()
        C = x_dbl[:, -d_state:].reshape(R, L, 1, d_state).permute(0, 2, 3, 1).contiguous()
        out, scan_intermediates, out_z = selective_scan_cuda.fwd(conv1d_out, delta, t["A"], B, C, t["D"], z, t["delta_bias"], True)
        y = F.linear(out_z.transpose(1, 2), t["out_proj_w"], None)
        y_nat = torch.gather(y, 1, rev[None, :, None].expand_as(y))
        assert rel_err(y_nat, c["out"]) <= 1e-5, rel_err(y_nat, c["out"])
        last_state = scan_intermediates[:, :, -1, 1::2]                     # selective_scan_interface.py:39
        assert last_state.shape == (R, twoD // 2, d_state) and torch.isfinite(last_state).all()

        # ---- backward, selective_scan_interface.py:901-1006 (checkpoint_lvl == 1)
        dout_nat = c["dout"].cuda()
        dout = torch.gather(dout_nat, 1, perm[None, :, None].expand_as(dout_nat)).contiguous()
        conv1d_out = causal_conv1d_cuda.causal_conv1d_fwd(x, conv1d_weight, t["conv_b"], True)
        delta = (t["dt_proj_w"] @ x_dbl[:, :delta_rank].t()).view(-1, R, L).transpose(0, 1)
        dxz = torch.empty_like(xz)
        dx, dz = dxz.chunk(2, dim=1)
        dout2 = dout.reshape(R * L, -1).t()
        dout_y = (t["out_proj_w"].t() @ dout2).view(-1, R, L).transpose(0, 1)
        dconv1d_out, ddelta, dA, dB, dC, dD, ddelta_bias, dz_ret, out_z2 = selective_scan_cuda.bwd(
            conv1d_out, delta, t["A"], B, C, t["D"], z, t["delta_bias"], dout_y, scan_intermediates, out, dz, True, True)
        assert dz_ret.data_ptr() == dz.data_ptr()                            # pre-allocated view of dxz is written in place
        assert rel_err(out_z2, out_z) <= 1e-6
        dout_proj_weight = torch.einsum("eB,dB->ed", dout2, out_z2.transpose(0, 1).reshape(twoD // 2, R * L))
        dx_dbl = torch.empty_like(x_dbl)
        dx_dbl[:, delta_rank:delta_rank + d_state] = dB.squeeze(1).transpose(1, 2).reshape(R * L, d_state)
        dx_dbl[:, -d_state:] = dC.squeeze(1).transpose(1, 2).reshape(R * L, d_state)
        ddelta2 = ddelta.transpose(0, 1).reshape(twoD // 2, R * L)
        ddelta_proj_weight = torch.einsum("dB,Br->dr", ddelta2, x_dbl[:, :delta_rank])
        dx_dbl[:, :delta_rank] = torch.einsum("dB,dr->Br", ddelta2, t["dt_proj_w"])
        dconv1d_out2 = dconv1d_out.transpose(0, 1).reshape(twoD // 2, R * L)
        dx_proj_weight = torch.einsum("Br,Bd->rd", dx_dbl, conv1d_out.transpose(1, 2).reshape(R * L, -1))
        dconv1d_out2 = torch.addmm(dconv1d_out2, t["x_proj_w"].t(), dx_dbl.t())
        dconv1d_out3 = dconv1d_out2.view(-1, R, L).transpose(0, 1)
        dconv1d_out3 = dconv1d_out3 if dconv1d_out3.stride(-1) == 1 else dconv1d_out3.contiguous()
        dx_ret, dconv1d_weight, dconv1d_bias = causal_conv1d_cuda.causal_conv1d_bwd(x, conv1d_weight, t["conv_b"], dconv1d_out3, dx, True)
        assert dx_ret.data_ptr() == dx.data_ptr()
        dxz_nat = torch.gather(dxz, 2, rev[None, None, :].expand_as(dxz))
        got = {"xz": dxz_nat, "conv_w": dconv1d_weight.view_as(t["conv_w"]), "conv_b": dconv1d_bias, "x_proj_w": dx_proj_weight,
               "dt_proj_w": ddelta_proj_weight, "out_proj_w": dout_proj_weight, "A": dA, "D": dD, "delta_bias": ddelta_bias}
        for n, g in got.items():
            assert rel_err(g, c["d" + n]) <= 2e-5, (n, rel_err(g, c["d" + n]))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def test_shims_raise_like_the_pybind_modules():
    """TORCH_CHECK failures of the reference (selective_scan.cpp:235-305, causal_conv1d.cpp:227-257) surface as RuntimeError,
    legal-but-unbuilt variants as NotImplementedError; nothing falls back."""
    import dimsum_b200.causal_conv1d_cuda as ccc
    import dimsum_b200.selective_scan_cuda as ssc
    u = torch.randn(2, 8, 16, device="cuda")
    A = -torch.rand(8, 4, device="cuda")
    Bm = torch.randn(2, 1, 4, 16, device="cuda")
    with pytest.raises(RuntimeError):
        ssc.fwd(u, u[:, :4], A, Bm, Bm, None, None, None, False)                       # delta shape
    with pytest.raises(RuntimeError):
        ssc.fwd(u, u, A.half(), Bm, Bm, None, None, None, False)                       # A dtype
    with pytest.raises(RuntimeError):
        ccc.causal_conv1d_fwd(u, torch.randn(8, 5, device="cuda"), None, True)         # width 5
    with pytest.raises(NotImplementedError):
        ccc.causal_conv1d_fwd(u.transpose(1, 2).contiguous().transpose(1, 2), torch.randn(8, 4, device="cuda"), None, True)
    with pytest.raises(NotImplementedError):
        ccc.causal_conv1d_update(u, u, A, None, True)
