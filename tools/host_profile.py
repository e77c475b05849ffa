"""cProfile of the host side of one fused-loss step on a tiny problem (GPU time negligible)."""
import cProfile
import math
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mmlearn_b200 import ContrastiveLoss, LossPairSpec

N, d = 2048, 512
dev = torch.device("cuda:0")
a = torch.randn(N, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
b = torch.randn(N, d, device=dev, dtype=torch.bfloat16, requires_grad=True)
ids = torch.stack([torch.zeros(N, dtype=torch.long), torch.arange(N)], 1).to(dev)
log_s = torch.tensor(math.log(1 / 0.07), device=dev, requires_grad=True)
fn = ContrastiveLoss(l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=True)
pairs = [LossPairSpec(("rgb", "text"))]


def step():
    a.grad = b.grad = log_s.grad = None
    loss = fn({"rgb_embedding": a, "text_embedding": b}, {"rgb": ids, "text": ids}, log_s.exp(), pairs)
    loss.backward()


for _ in range(20):
    step()
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    step()
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
