#!/usr/bin/env python
"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: isolate ONE train step
(the launches between two consecutive adam_kernel launches) and rank kernels by device time.
    python tools/launch_summary.py gpurun_out/launches.csv [step_index_from_end]"""
import collections
import csv
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    out = []
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)
        out.append((row["Kernel Name"], v))
    return out


def main():
    rows = load(sys.argv[1])
    back = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    ends = [i for i, (k, _) in enumerate(rows) if "adam_kernel" in k]
    if len(ends) < back + 1:
        print("not enough steps in the capture", len(ends))
        return
    lo, hi = ends[-back - 1] + 1, ends[-back] + 1
    step = rows[lo:hi]
    agg = collections.OrderedDict()
    for k, v in step:
        k = k.split("(")[0][:90]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "bvae::" in k)
    print(f"one step: {len(step)} launches, {tot:.1f} us device time ({ours:.1f} us in bvae:: kernels, "
          f"{tot - ours:.1f} us in ATen/cuFFT kernels)")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{a[1]:9.1f} us {100 * a[1] / tot:5.1f}%  {a[0]:3d} x {a[1] / a[0]:8.1f}  {k}")


if __name__ == "__main__":
    main()
