"""Retrieval recall@k: the fused metric on the GPU against the
reference's arithmetic (CPU matmul + topk in a thread pool, retrieval_recall.py:199-223) on the
host, same synthetic embeddings.  The CPU leg runs on a bounded sample of the queries against the
whole database (same per-query work) and is scaled to queries/s.
usage: python tools/recall_bench.py [n] [d] [cpu_queries]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mmlearn_b200 import RetrievalRecallAtK

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
cpu_q = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
x = torch.randn(n, d, generator=g).to(torch.bfloat16).float()
y = (0.2 * x + torch.randn(n, d, generator=g)).to(torch.bfloat16).float()
idx = torch.arange(n)
xd, yd, idd = x.to(dev), y.to(dev), idx.to(dev)
bs = 4096


def gpu_once():
    met = RetrievalRecallAtK(top_k=10, reduction="none")
    for i in range(0, n, bs):
        met.update(xd[i:i + bs], yd[i:i + bs], torch.arange(min(bs, n - i), device=dev))
    return [float(met.compute(k)) for k in (1, 5, 10)]


for _ in range(2):
    vals = gpu_once()
torch.cuda.synchronize()
t0 = time.perf_counter()
reps = 5
for _ in range(reps):
    vals = gpu_once()
torch.cuda.synchronize()
t_gpu = (time.perf_counter() - t0) / reps

# reference arithmetic on the host: normalise, x_batch @ y.T, topk, gather (one k; the reference
# repeats all of it for every k)
torch.set_num_threads(os.cpu_count())
xn = x / x.norm(dim=-1, keepdim=True)
yn = y / y.norm(dim=-1, keepdim=True)
t0 = time.perf_counter()
sim = xn[:cpu_q] @ yn.T
top = torch.topk(sim, 10, dim=1)[1]
hit = (top == idx[:cpu_q, None]).any(1)
t_cpu = time.perf_counter() - t0
ref10 = float(hit.float().mean())
print(f"recall@k n={n} d={d}: fused GPU metric (update + compute, k = 1, 5, 10 from one pass) "
      f"{t_gpu * 1e3:.2f} ms = {n / t_gpu / 1e6:.2f} M queries/s; R@1/5/10 = {vals}")
print(f"reference arithmetic on the host ({os.cpu_count()} threads, {cpu_q} queries x {n} db, one k): "
      f"{t_cpu * 1e3:.1f} ms = {cpu_q / t_cpu / 1e3:.1f} k queries/s; R@10 on the sample = {ref10:.4f}")
