"""Per-kernel SASS mnemonic histogram of libdimsum_b200.so (cuobjdump -sass), the evidence for which hardware paths a kernel
uses: UTCHMMA / UTCQMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG (TMA tensor loads), UBLKCP (bulk copies), SYNCS
(mbarrier), LDGSTS (cp.async), FFMA2 / FMUL2 (packed fp32), MUFU.*.

    python tools/sass_summary.py > profiles/r2_sass_summary.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dimsum_b200", "libdimsum_b200.so")
KEEP = ("UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "SYNCS", "UTCBAR", "UTCATOMSWS", "LDGSTS", "FFMA2",
        "FMUL2", "FADD2", "MUFU", "HMMA", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "ATOMG", "RED", "FFMA", "F2FP")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    kernels = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            kernels[cur][m.group(1)] += 1
    names = list(kernels)
    dm = subprocess.run(["cu++filt"] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines() if names else []
    for n, d in zip(names, dm):
        demangle[n] = d
    print("# SASS summary of `dimsum_b200/libdimsum_b200.so` (sm_100a) -- `python tools/sass_summary.py`\n")
    print("Instruction counts per kernel (static SASS, `cuobjdump -sass`), grouped by mnemonic prefix; only the kernels and prefixes that "
          "matter for the hardware-path evidence are listed.  `UTCHMMA` = tcgen05.mma, `LDTM` / `STTM` = tcgen05.ld / st, `UTMALDG` = "
          "TMA tensor load, `UBLKCP` = bulk copy, `SYNCS` = mbarrier, `LDGSTS` = cp.async, `FFMA2` / `FMUL2` = packed fp32.\n")
    tot = collections.Counter()
    rows = []
    for n, c in kernels.items():
        grouped = collections.Counter()
        for op, k in c.items():
            for pre in KEEP:
                if op == pre or op.startswith(pre + "."):
                    grouped[pre] += k
                    break
        tot.update(grouped)
        short = re.sub(r"\(.*", "", demangle.get(n, n))
        short = short.replace("dimsum::(anonymous namespace)::", "").replace("void ", "")
        rows.append((short, sum(c.values()), grouped))
    cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "SYNCS", "LDGSTS", "FFMA2", "FMUL2", "MUFU", "LDS", "STS", "SHFL", "ATOMG", "RED"]
    print("| kernel | SASS instr | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    seen = set()
    for short, n, g in rows:
        if short in seen and not any(g[c] for c in ("UTCHMMA", "UTMALDG")):
            continue
        seen.add(short)
        print(f"| `{short[:90]}` | {n} | " + " | ".join(str(g[c]) if g[c] else "" for c in cols) + " |")
    print("\nTotals over the library: " + ", ".join(f"{k} {v}" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if k in cols))


if __name__ == "__main__":
    sys.exit(main())
