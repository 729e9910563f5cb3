#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native kernel (B200_PROFILING.md):
UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG / UBLKCP (TMA), LDTM / STTM (tcgen05.ld / st), UTCBAR (tcgen05.commit),
LDGMC / multimem (NVLS), HMMA (warp-level mma.sync).      python tools/sass_summary.py > profiles/sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "beta-vae_b200", "libbvae_b200.so")
PATS = ["UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "LDGMC", "STG.*MC|REDG.*MC|ST.*MULTIMEM", "HMMA", "ELECT",
        "BRA.U.ANY"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        for p in PATS:
            if re.search(r"\b(" + p + r")\b", line) or (p in ("UTCHMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "LDGMC", "HMMA")
                                                          and re.search(r"\s" + p + r"[\.\s]", line)):
                counts[cur][p] += 1
    names = list(counts)
    dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
    print(f"# {os.path.relpath(LIB, ROOT)}: SASS mnemonic counts per kernel (cuobjdump -sass); kernels without any of them omitted")
    print("# " + "  ".join(f"{p[:10]:>10s}" for p in PATS) + "  kernel")
    tot = collections.Counter()
    for n, d in zip(names, dem):
        c = counts[n]
        if not any(c[p] for p in PATS[:10]):
            continue
        tot.update(c)
        short = re.sub(r"\(.*", "", d).replace("void bvae::", "").replace("bvae::", "")
        print("  " + "  ".join(f"{c[p]:10d}" for p in PATS) + "  " + short)
    print("  " + "  ".join(f"{tot[p]:10d}" for p in PATS) + "  TOTAL")
    print("# BRA.U.ANY = waterfall loops around uniform-datapath instructions; the tcgen05 kernels have none since the "
          "producer / issuer loops became warp-uniform (round 2)")


if __name__ == "__main__":
    sys.exit(main())
