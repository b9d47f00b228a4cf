"""GPU parity of the fused wavelet-packet / window-scan kernels and the token gather."""
import numpy as np
import pytest
import torch

from golden_io import load, rel_err, tol

pytestmark = pytest.mark.gpu
GOLD = load("wavelet.npz")


@pytest.mark.parametrize("case", ["g16_c32", "g32_c16", "g8_c48"])
def test_packet_matches_reference_golden(case):
    from dimsum_b200 import wavelet_packet, wavelet_packet_inverse
    c = GOLD[case]
    assert rel_err(wavelet_packet(c["x"].cuda()), c["coef"]) <= 1e-5
    assert rel_err(wavelet_packet_inverse(c["c"].cuda()), c["recon"]) <= 1e-5


def test_index_map_is_bit_exact():
    """Feed one-hot-free integer-valued images: every output is an exact small dyadic rational, so any wrong
    token/channel index or sign shows up as a non-zero difference (the 'bit-exact wavelet index tables' claim)."""
    from dimsum_b200 import wavelet_packet, wavelet_packet_inverse
    from oracle import ref_ops
    g = torch.Generator().manual_seed(5)
    for grid, C in ((16, 512), (32, 64), (8, 16)):
        x = torch.randint(-8, 9, (2, grid * grid, C), generator=g).float() * 16.0
        assert torch.equal(wavelet_packet(x.cuda()).cpu(), ref_ops.wavelet_packet_oracle(x))
        assert torch.equal(wavelet_packet_inverse(x.cuda()).cpu(), ref_ops.wavelet_packet_inverse_oracle(x))


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("column_first", [False, True])
def test_fused_window_scan_round_trip_full_size(dtype, column_first):
    """Model shape (512 rows, 16x16 tokens, 512 channels): dwt+local_scan fused, idwt+local_reverse fused, round trip."""
    from dimsum_b200 import scanning_orders as so, wavelet_packet, wavelet_packet_inverse
    from oracle import ref_ops
    grid, C, R = 16, 512, 512
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(R, grid * grid, C, generator=g, device="cuda").to(dtype)
    order = so.window_order(grid, grid // 4, column_first)
    pos = so.as_index(so.reverse_permut_np(order), "cuda")
    y = wavelet_packet(x, pos)
    want = ref_ops.window_scan_oracle(ref_ops.wavelet_packet_oracle(x[:3].float().cpu()), grid // 4, column_first)
    assert rel_err(y[:3], want) <= tol(dtype)
    back = wavelet_packet_inverse(y, pos)
    assert rel_err(back, x) <= (2e-6 if dtype == torch.float32 else 2e-2)


def test_gradients_are_the_adjoint():
    from dimsum_b200 import wavelet_packet, wavelet_packet_inverse
    c = GOLD["grad_g8_c16"]
    x = c["x"].cuda().requires_grad_(True)
    (gx,) = torch.autograd.grad(wavelet_packet(x), x, c["gy"].cuda())
    assert rel_err(gx, c["gx"]) <= 1e-5
    y = torch.randn(2, 64, 32, device="cuda", requires_grad=True)
    g = torch.randn(2, 64, 32, device="cuda")
    (gy,) = torch.autograd.grad(wavelet_packet_inverse(y), y, g)
    assert rel_err(gy, 16.0 * wavelet_packet(g)) <= 1e-5


def test_token_gather_and_local_scan_match_tables():
    from dimsum_b200 import scanning_orders as so
    orders = np.load(__import__("os").path.join(__import__("golden_io").GOLDEN, "orders.npz"))
    x = torch.randn(3, 256, 64, device="cuda")
    for cf in (0, 1):
        seq = torch.from_numpy(orders[f"window_16_4_{cf}"].astype(np.int64)).cuda()
        y = so.local_scan(x, w=4, H=16, W=16, column_first=bool(cf))
        assert torch.equal(y, x[:, seq])
        assert torch.equal(so.local_reverse(y, w=4, H=16, W=16, column_first=bool(cf)), x)
    for dtype in (torch.float32, torch.bfloat16):
        xx = torch.randn(2, 1024, 512, device="cuda").to(dtype)
        perm = so.zigma_path(32)[3]
        assert torch.equal(so.token_gather(xx, so.as_index(perm, "cuda")), xx[:, torch.from_numpy(perm).cuda()])
    xg = torch.randn(2, 256, 32, device="cuda", requires_grad=True)
    so.local_scan(xg, w=4, H=16, W=16).square().sum().backward()
    assert rel_err(xg.grad, 2 * xg.detach()) <= 1e-6
