"""Top stall-sampled SASS instructions of an `ncu --page source --csv` dump.
usage: python tools/ncu_top.py dump.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ia, isamp = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
data = [(int(r[isamp] or 0), idx, r[ia]) for idx, r in enumerate(rows[h + 1:]) if len(r) > isamp]
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, idx, src in sorted(data, reverse=True)[:n]:
    print(f"{s:7d} {100 * s / tot:5.1f}% @{idx:5d} {src.strip()[:110]}")
