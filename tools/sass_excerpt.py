"""Blackwell-native evidence from the built library: per kernel, the count of the SASS mnemonics
that tcgen05 / TMEM / TMA / NVLink-peer code compiles to, plus the first lines that carry each of
them.  usage: python tools/sass_excerpt.py [libmmcl.so] > profiles/r02_sass_excerpts.md"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(
    os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mmlearn_b200", "libmmcl.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
WANT = OrderedDict([
    ("UTCHMMA", "tcgen05.mma kind::f16 (bf16 x bf16 -> fp32 in tensor memory)"),
    ("UTCBAR", "tcgen05.commit -> mbarrier"),
    ("LDTM", "tcgen05.ld (tensor memory -> registers)"),
    ("STTM", "tcgen05.st (registers -> tensor memory: bf16 G as the A operand of dX += G Y)"),
    ("UTMALDG", "cp.async.bulk.tensor (TMA tile loads)"),
    ("SYNCS", "mbarrier arrive / try_wait"),
    ("MUFU.EX2", "ex2.approx (softmax / sigmoid exponentials)"),
    ("FFMA2", "packed fp32x2 FMA"),
    ("REDG", "red.global (fp32x4 reductions of split segments, ABI accumulate mode)"),
    ("HMMA", "legacy mma.sync path (must be absent)"),
])
KERNELS = ["sim_lse_fwd_kernelILb1ELi0", "sim_grad_bwd_pair_kernel", "sim_grad_bwd_kernelILb1",
           "step_post_fwd_kernel", "step_pack_kernel", "step_unpack_kernelILi2"]
blocks = re.split(r"\n\s*Function : ", sass)
print("# SASS evidence (cuobjdump -sass mmlearn_b200/libmmcl.so, sm_100a)\n")
print("Mnemonic counts per kernel; an excerpt of the first occurrences follows each table.\n")
for want in KERNELS:
    for b in blocks[1:]:
        name = b.split("\n", 1)[0].strip()
        if want not in name:
            continue
        lines = [ln for ln in b.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/\s+[A-Z@]", ln)]
        print(f"## `{name[:110]}`\n")
        print(f"{len(lines)} SASS instructions\n")
        print("| mnemonic | count | what it is |")
        print("|---|---|---|")
        ex = []
        for m, what in WANT.items():
            hits = [ln.strip() for ln in lines if re.search(r"(?<![A-Z])" + re.escape(m), ln)]
            print(f"| `{m}` | {len(hits)} | {what} |")
            ex += hits[:2]
        peer = [ln.strip() for ln in lines if re.search(r"ST\.E\.(128|64)\.STRONG\.SYS|STG\.E\.128|ST\.E\.128", ln)]
        sysld = [ln.strip() for ln in lines if "LD.E.64.STRONG.SYS" in ln or "LDG.E.64.STRONG.SYS" in ln]
        print(f"| 16-byte global stores | {len(peer)} | in sim_lse_fwd / step_pack these include the peer-arena pushes (plain stores: the addresses are CUDA-IPC mappings of the other GPUs' arenas) |")
        print(f"| `LD*.STRONG.SYS` 64-bit | {len(sysld)} | ld.acquire.sys on the per-source ready flags |")
        print("\n```")
        for ln in ex[:14] + sysld[:2]:
            print(re.sub(r"\s+", " ", ln)[:150])
        print("```\n")
        break
