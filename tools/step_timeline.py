"""Kernel timeline of a few fused-loss steps (torch.profiler / CUPTI): per-kernel start, duration
and the idle gap before it.  usage: python tools/step_timeline.py [N] [d] [steps]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile

from mmlearn_b200 import ContrastiveLoss, LossPairSpec

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
nosync = len(sys.argv) > 4 and sys.argv[4] == "nosync"  # let the host run ahead between steps
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
nb = N // world  # rows of this rank (global batch N)
g = torch.Generator().manual_seed(1234 + rank)
a = torch.randn(nb, d, generator=g).to(torch.bfloat16).to(dev).requires_grad_(True)
b = (0.5 * a.detach().float().cpu() + torch.randn(nb, d, generator=g)).to(torch.bfloat16).to(dev).requires_grad_(True)
ids = torch.stack([torch.zeros(nb, dtype=torch.long), torch.arange(rank * nb, (rank + 1) * nb)], 1).to(dev)
log_s = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)
fn = ContrastiveLoss(l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=True)
pairs = [LossPairSpec(("rgb", "text"))]


def step():
    a.grad = b.grad = log_s.grad = None
    loss = fn({"rgb_embedding": a, "text_embedding": b}, {"rgb": ids, "text": ids}, log_s.exp(), pairs)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        step()
        if not nosync:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
if rank != 0:
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.exit(0)
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
prev_end = t0
tot_busy = 0.0
for e in evs:
    st, en = e.time_range.start, e.time_range.end
    print(f"{(st - t0):10.1f} us  dur {(en - st):9.1f}  gap {(st - prev_end):8.1f}  {e.name[:70]}")
    prev_end = max(prev_end, en)
    tot_busy += en - st
print(f"span {(prev_end - t0):.1f} us, busy {tot_busy:.1f} us over {steps} steps")
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
