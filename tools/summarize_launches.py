"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of device time per kernel.

    python tools/summarize_launches.py gpurun_out/launches.csv [top_n] > profiles/launches_summary.md
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    r = csv.reader(lines)
    hdr = next(r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    n = 0
    for row in r:
        if len(row) <= vi:
            continue
        v = float(row[vi].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(row[ui], 1.0)
        short = re.sub(r"\(.*", "", row[ki])
        short = re.sub(r"void |dimsum::<unnamed>::|at::native::|<unnamed>::", "", short)[:100]
        agg[short][0] += 1
        agg[short][1] += v
        n += 1
    tot = sum(v[1] for v in agg.values())
    mine = ("scan_fwd", "scan_bwd", "conv_fwd", "conv_bwd", "conv_xproj", "attention_kernel", "wavelet_kernel", "gather_kernel",
            "rowwise_kernel", "add_rmsnorm", "norm_kernel", "gelu_mul", "colsum", "rmsnorm_bwd")
    ours = sum(v[1] for k, v in agg.items() if "dimsum" in k or any(t in k for t in mine))
    print(f"launches: {n}   total device time: {tot / 1e3:.1f} ms   this repo's kernels: {100 * ours / tot:.1f} % of device time\n")
    print("| share | launches | avg us | kernel |")
    print("|---:|---:|---:|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"| {100 * v[1] / tot:.2f} % | {v[0]} | {v[1] / v[0]:.1f} | `{k}` |")


if __name__ == "__main__":
    main()
