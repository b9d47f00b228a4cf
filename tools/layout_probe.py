"""Developer probe: inference scan time vs. where u / delta / z live in memory (same kernel, same bytes).

    python tools/layout_probe.py

The mixer hands the scan row-strided views of GEMM outputs; this measures how much the relative placement of the three
streams matters on the HBM side, to choose the layout the mixer should produce.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tools.bench_ops import timeit  # noqa: E402


def shifted(shape_fn, numel, pad, dtype):
    """A tensor built by shape_fn on a fresh buffer whose first element sits `pad` elements into the allocation."""
    buf = torch.empty(numel + pad, dtype=dtype, device="cuda")
    return shape_fn(buf[pad:pad + numel]), buf


def main():
    from dimsum_b200 import selective_scan_cuda
    R, D, L, N = 512, 1024, 256, 16
    dtype = torch.float32
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    A = -0.5 * torch.rand(D, N, generator=g, device="cuda") - 0.05
    Bm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
    Cm = torch.randn(R, 1, N, L, generator=g, device="cuda").to(dtype)
    Dv = torch.ones(D, device="cuda")
    bias = torch.rand(D, device="cuda") - 3.0

    def run(name, u, delta, z):
        u.normal_(); z.normal_(); delta.uniform_(0, 0.5)
        med, best = timeit(lambda: selective_scan_cuda.fwd(u, delta, A, Bm, Cm, Dv, z, bias, True, need_out=False, need_x=False),
                           flush=flush)
        print(f"{name:60s} {med:7.3f} ms (best {best:.3f})   u@{u.data_ptr() % (1 << 30):#x} d@{delta.data_ptr() % (1 << 30):#x} "
              f"z@{z.data_ptr() % (1 << 30):#x}", flush=True)

    n = R * D * L
    rows_major = lambda t: t.view(R, D, L)                          # (batch, channel, L) contiguous
    chan_major = lambda t: t.view(D, R, L).transpose(0, 1)          # view of a (channels, batch*L) GEMM output
    for pad_d, pad_z in [(0, 0), (64, 128), (1024, 2048), (16384 + 64, 32768 + 128), ((1 << 18) + 1024, (1 << 19) + 2048)]:
        u, _ = shifted(rows_major, n, 0, dtype)
        d, _b1 = shifted(rows_major, n, pad_d, dtype)
        z, _b2 = shifted(rows_major, n, pad_z, dtype)
        run(f"all (b,d,L) contiguous, delta +{pad_d * 4} B, z +{pad_z * 4} B", u, d, z)
        del u, d, z, _b1, _b2
    for pad_d, pad_z in [(0, 0), (64, 128), (1024, 2048), (16384 + 64, 32768 + 128), ((1 << 18) + 1024, (1 << 19) + 2048)]:
        u, _ = shifted(rows_major, n, 0, dtype)
        d, _b1 = shifted(chan_major, n, pad_d, dtype)
        z, _b2 = shifted(chan_major, n, pad_z, dtype)
        run(f"mixer: u contiguous, delta/z channel-major, delta +{pad_d * 4} B, z +{pad_z * 4} B", u, d, z)
        del u, d, z, _b1, _b2
    xz = torch.empty(R, 2 * D, L, dtype=dtype, device="cuda")
    d, _b = shifted(chan_major, n, 0, dtype)
    run("u, z halves of one (b, 2d, L) buffer, delta channel-major", xz[:, :D], d, xz[:, D:])
    xzt = torch.empty(2 * D, R, L, dtype=dtype, device="cuda").transpose(0, 1)
    run("u, z halves of one channel-major buffer, delta channel-major", xzt[:, :D], d, xzt[:, D:])
    u, _ = shifted(chan_major, n, 0, dtype)
    run("all three channel-major (own buffers)", u, d, xzt[:, D:])


if __name__ == "__main__":
    main()
