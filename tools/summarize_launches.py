#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into a per-iteration kernel table.
usage: tools/summarize_launches.py launches.csv [iteration index] > profiles/<name>.md"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [(r["Kernel Name"], float(r["Metric Value"])) for r in csv.DictReader(lines)]
    marks = [i for i, (k, _) in enumerate(rows) if "grid_sample_fwd_kernel" in k or "grid_sample_fwd_packed_kernel" in k]
    a, b = marks[which], marks[which + 1]
    it = rows[a:b]
    tot = sum(v for _, v in it)
    agg = collections.OrderedDict()
    for k, v in it:
        k2 = re.sub(r"\(.*", "", k).replace("void ", "").replace("<unnamed>::", "")[:100]
        agg.setdefault(k2, [0, 0.0])
        agg[k2][0] += 1
        agg[k2][1] += v
    ours = sum(v for k, v in it if "<unnamed>::" in k and "at::" not in k)
    print(f"# ncu launch list: one warm attack iteration (launches {a}..{b} of `{path}`)\n")
    print("`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, not absolutes)\n")
    print(f"* launches in the iteration: {len(it)}; summed kernel time {tot / 1e6:.3f} ms")
    print(f"* share of spaa_b200 kernels: {100 * ours / tot:.1f}% (the rest is the external cuDNN/ATen classifier)\n")
    print("| time (us) | share | launches | kernel |\n|---:|---:|---:|---|")
    for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
        print(f"| {v / 1e3:.1f} | {100 * v / tot:.1f}% | {n} | `{k}` |")


if __name__ == "__main__":
    main()
