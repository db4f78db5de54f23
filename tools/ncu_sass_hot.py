"""Hot SASS lines of one launch in an ncu report's source page.
usage: ncu -i rep --page source --csv --print-source sass --launch-skip K --launch-count 1 > x.csv; python tools/ncu_sass_hot.py x.csv [frac]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
a, b = starts[0], starts[1]
print(rows[a][1][:120])
hdr = rows[a + 1]
data = [r for r in rows[a + 2:b] if len(r) == len(hdr)]
ia, isamp = hdr.index("Instructions Executed"), hdr.index("# Samples")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[ia]) for r in data); ts = sum(int(r[isamp]) for r in data)
print("warp instructions", tot, "samples", ts, "SASS lines", len(data))
agg = {hdr[i]: sum(int(r[i]) for r in data) for i in stalls}
print("stalls:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > ts * 0.01})
for i, r in enumerate(data):
    if int(r[ia]) > tot * frac or int(r[isamp]) > ts * frac * 1.5:
        st = sorted(((int(r[j]), hdr[j][6:]) for j in stalls), reverse=True)[:2]
        print(f"{i:5d} ex={int(r[ia]):8d} smp={int(r[isamp]):5d} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]} | {r[1].strip()[:90]}")
