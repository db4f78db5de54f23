#!/usr/bin/env python
"""Top stall sites per kernel from `ncu -i rep --page source --csv --print-source sass` output.
usage: python tools/ncu_hot.py sass.csv [kernel index | -1 for the list] [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
which = int(sys.argv[2]) if len(sys.argv) > 2 else -1
n = int(sys.argv[3]) if len(sys.argv) > 3 else 25
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
if which < 0:
    for k, i in enumerate(starts):
        print(k, rows[i][1][:90])
    sys.exit(0)
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[lo][1][:100])
hdr = rows[lo + 1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[lo + 2:hi] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
order = sorted(range(len(body)), key=lambda i: -int(body[i][ix["# Samples"]]))[:n]
for i in sorted(order):
    r = body[i]
    top = sorted(((int(r[ix[s]]), s) for s in stalls), reverse=True)[:2]
    print(f"{i:6d} {int(r[ix['# Samples']]):6d} {100.0 * int(r[ix['# Samples']]) / tot:5.1f}%  {r[ix['Source']].strip()[:80]:80s} {top}")
