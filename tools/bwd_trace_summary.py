"""Phase durations of the pair backward kernel from tools/bwd_trace's CSV (SM clocks, one CTA pair).
usage: python tools/bwd_trace_summary.py trace.csv"""
import csv
import statistics
import sys
from collections import defaultdict

NAMES = {0: "S.issue", 1: "S.issued", 2: "dXp.wait", 3: "dXp.go", 4: "dXp.issued", 5: "dXo.q0",
         6: "dXo.q2", 7: "dXo.issued", 8: "prod.S0", 9: "prod.S3", 10: "epi0.S", 11: "epi1.S",
         12: "epi0.free", 13: "epi1.free", 14: "epi0.done", 15: "epi1.done"}
NAMES.update({16 + k: f"S.data{k}" for k in range(8)})
ev = defaultdict(dict)  # (cta, step) -> {event: clk}
for r in csv.DictReader(open(sys.argv[1])):
    ev[(int(r["cta"]), int(r["step"]))][int(r["event"])] = int(r["clk"])


def med(pairs, label):
    v = [b - a for a, b in pairs]
    if v:
        q = statistics.quantiles(v, n=10)
        print(f"  {label:46s} median {statistics.median(v):8.0f}  p10 {q[0]:8.0f}  p90 {q[-1]:8.0f}  n={len(v)}")


for c in (0, 1):
    steps = sorted(s for (cc, s) in ev if cc == c)
    steps = [s for s in steps if 10 <= s < max(steps) - 10]  # steady state only
    get = lambda s, e: ev.get((c, s), {}).get(e)
    span = lambda a, b, ds=0: [(get(s, a), get(s + ds, b)) for s in steps
                               if get(s, a) is not None and get(s + ds, b) is not None]
    print(f"CTA {c}: {len(steps)} steady steps (clk)")
    med(span(0, 0, 1), "period: S.issue(k) -> S.issue(k+1)")
    med(span(0, 10), "S: issue start -> epilogue sees S")
    med(span(0, 16), "   issue start -> stage 0 data")
    for k in range(1, 8):
        med(span(15 + k, 16 + k), f"   stage {k - 1} data -> stage {k} data")
    med(span(23, 1), "   stage 7 data -> all issued")
    med(span(1, 10), "   all issued -> epilogue sees S")
    med(span(10, 12), "epilogue: S seen -> partner buffer free")
    med(span(12, 14), "conversion (wg0): free -> done")
    med(span(13, 15), "conversion (wg1): free -> done")
    med(span(14, 0, 1), "conv done (wg0) -> next S.issue")
    med(span(15, 0, 1), "conv done (wg1) -> next S.issue")
    med(span(7, 0, 1), "dX own issued -> next S.issue")
    med(span(1, 2), "S issued -> dX partner wait begins")
    med(span(2, 3), "dX partner: wait for G tile")
    med(span(3, 4), "dX partner: issue")
    med(span(10, 5), "S seen -> quarter 0 handed over")
    med(span(5, 6), "quarter 0 -> quarter 2 handed over")
    med(span(6, 7), "quarter 2 -> dX own all issued")
    med(span(8, 9), "producer: S stage 0 slot -> S stage 3 slot")
    med(span(8, 0), "producer S stage 0 slot -> S.issue (lead)")
    if any(e >= 32 for (cc, s_), d in ev.items() if cc == c for e in d):
        print("  conversion loop of the first epilogue warp (per 32-column chunk ci):")
        for ci in range(4):
            b = 32 + ci * 8
            prev = (12 if ci == 0 else 32 + (ci - 1) * 8 + 5)
            med(span(prev, b), f"    ci={ci}: previous -> first half loaded")
            med(span(b, b + 1), f"    ci={ci}: convert first 16 columns")
            med(span(b + 1, b + 2), f"    ci={ci}: ld wait + convert second 16")
            med(span(b + 2, b + 3), f"    ci={ci}: tmem store + transposition stores")
            med(span(b + 3, b + 4), f"    ci={ci}: transposed loads + st.async")
            med(span(b + 4, b + 5), f"    ci={ci}: tmem store wait + fence + arrive")
