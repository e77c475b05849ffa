"""Summary of an `ncu --set full` capture: per launch, duration, DRAM bytes, tensor-pipe activity,
L2 -> SM bytes, registers, clock.  Also writes the DRAM traffic per (kernel, shape) that bench.py
reports as `roofline.traffic` (profiles/ncu_traffic.json).
usage: python tools/ncu_summary.py capture.ncu-rep out.md [nx ny d]"""
import csv
import io
import json
import os
import subprocess
import sys

rep, out_md = sys.argv[1], sys.argv[2]
shape = [int(v) for v in sys.argv[3:6]] if len(sys.argv) >= 6 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True,
                     check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
M = [("gpu__time_duration.sum", "time"),
     ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
     ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
     ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM"),
     ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
     ("launch__registers_per_thread", "regs/thread"),
     ("launch__grid_size", "grid"), ("launch__block_size", "block"),
     ("sm__cycles_elapsed.avg.per_second", "SM clock")]


def to_bytes(v, u):
    f = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    return float(v) * f.get(u, 1)


lines = ["| kernel | " + " | ".join(n for _, n in M) + " |", "|---|" + "---|" * len(M)]
traffic = []
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = name.split("(")[0].split("::")[-1][:40]
    cells = []
    for key, _ in M:
        if key not in col:
            cells.append("-")
            continue
        v, u = r[col[key]], units[col[key]]
        try:
            cells.append(f"{float(v):.4g} {u}".strip())
        except ValueError:
            cells.append(v)
    lines.append(f"| `{short}` | " + " | ".join(cells) + " |")
    if "sim_grad_bwd_pair" in name and shape:
        b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
            to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
        traffic.append(b)
open(out_md, "w").write(
    f"# ncu --set full summary of `{os.path.basename(rep)}`\n\n"
    "One steady-state step of `python bench.py --steps 1 --warmup 3` (config 2, one GPU): the five\n"
    "launches of the fused step.  Per-launch times are cold-cache and serialised under the profiler:\n"
    "compare shares, not absolutes.\n\n" + "\n".join(lines) + "\n")
if traffic and shape:
    p = os.path.join(os.path.dirname(os.path.abspath(out_md)), "ncu_traffic.json")
    data = json.load(open(p)) if os.path.exists(p) else []
    data = [d for d in data if not (d["kernel"] == "sim_grad_bwd_pair_kernel" and
                                    (d["nx"], d["ny"], d["d"]) == tuple(shape))]
    # the fused launch covers both sides of the pair: report per side, like the timed call
    data.append({"kernel": "sim_grad_bwd_pair_kernel", "nx": shape[0], "ny": shape[1], "d": shape[2],
                 "dram_bytes": sum(traffic) / len(traffic) / 2,
                 "source": f"profiles/{os.path.basename(out_md)} (fused launch of both sides, halved)"})
    json.dump(data, open(p, "w"), indent=1)
print("\n".join(lines))
