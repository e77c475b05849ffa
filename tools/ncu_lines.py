"""Stall samples per CUDA source line from `ncu -i rep --page source --print-source cuda,sass --csv`.
usage: python tools/ncu_lines.py dump.csv [N]      (top N lines by samples, with the two main reasons)"""
import csv
import os
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40


def num(x):
    return int(x) if x.strip().isdigit() else 0


lines, fname, hdr = [], "?", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = os.path.basename(r[1])
    elif r[0] == "Line No":
        hdr = r
        isamp = hdr.index("Warp Stall Sampling (All Samples)")
        stall = [(i, c[6:]) for i, c in enumerate(hdr)
                 if c.startswith("stall_") and "Not Issued" not in c]
    elif hdr and r[0].isdigit() and len(r) > isamp:  # SASS rows have no line number
        s = num(r[isamp])
        top = sorted(((num(r[i]), c) for i, c in stall), reverse=True)[:2]
        lines.append((s, fname, int(r[0]), r[1].strip(), " ".join(f"{c}:{v}" for v, c in top if v)))
tot = sum(x[0] for x in lines)
print("total samples", tot)
for s, f, ln, src, why in sorted(lines, reverse=True)[:n]:
    print(f"{s:7d} {100 * s / tot:5.1f}% {f}:{ln:<5d} {src[:80]:80s} {why}")
