"""TEST / BASELINE INFRASTRUCTURE ONLY -- the reference's CPU arithmetic for ONE rank of the
headline workload (local_loss=True, gather_with_grad=True), used by ``bench.py`` for its
``cpu_baseline`` leg and its ``--impl reference`` arm.  The real mmlearn package cannot travel
to the GPU box (no wheel for its dependencies, /root/reference is absent there), so this is the
pinned port: every line restates mmlearn/modules/losses/contrastive.py.

A full step at N=32768 costs 6.6e12 FLOP plus a 2.1 GB id-match temporary on the CPU; the
bounded sample is the share of one rank holding ``b`` of the N rows: the same per-sample work
(12*N*d FLOP per sample, all N keys) on fewer samples.
"""

from __future__ import annotations

import time

import torch
import torch.nn.functional as F


def rank_share_step(a_all: torch.Tensor, b_all: torch.Tensor, b: int, scale: float):
    """fwd+bwd of rank 0's loss over its first ``b`` rows against all N gathered keys.

    ``a_all``/``b_all``: fp32 [N, d] raw embeddings (bf16-representable values), on the CPU or
    (``time_rank_share(device="cuda")``: the same torch ops through cuBLAS / ATen) on a GPU."""
    a_loc = a_all[:b].clone().requires_grad_(True)
    b_loc = b_all[:b].clone().requires_grad_(True)
    s = torch.tensor(scale, requires_grad=True, device=a_all.device)
    # contrastive.py:89-90 (every rank normalises its own rows; the gathered rows arrive normalised)
    fa = F.normalize(a_loc, p=2, dim=-1)
    fb = F.normalize(b_loc, p=2, dim=-1)
    with torch.no_grad():
        ga = F.normalize(a_all, p=2, dim=-1)
        gb = F.normalize(b_all, p=2, dim=-1)
    # gather_with_grad: the local rows sit inside the gathered tensors with grad (contrastive.py:506)
    all_a = torch.cat([fa, ga[b:]], 0)
    all_b = torch.cat([fb, gb[b:]], 0)
    # contrastive.py:327-332
    logits_a = s * (fa @ all_b.T)
    logits_b = s * (fb @ all_a.T)
    labels = torch.arange(b, device=a_all.device)  # contrastive.py:194-212 with off_0 = 0
    loss = (F.cross_entropy(logits_a, labels) + F.cross_entropy(logits_b, labels)) / 2
    loss.backward()
    return float(loss), a_loc.grad, b_loc.grad, float(s.grad)


def time_rank_share(n: int, d: int, b: int, steps: int, warmup: int, seed: int = 1234,
                    threads: int | None = None, device: str = "cpu", autocast_bf16: bool = False):
    """Median-free simple timing: total wall time of ``steps`` steps -> samples/s = b*steps/t.

    ``device="cuda"`` times the same restatement through torch's CUDA kernels (cuBLAS + ATen),
    optionally under bf16 autocast as Lightning's ``bf16-mixed`` would run the reference
    (mmlearn/cli/run.py:46-49): the library baseline on the same GPU."""
    if threads:
        torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    a = torch.randn(n, d, generator=g).to(torch.bfloat16).float()
    g = torch.Generator().manual_seed(seed + 1)
    bb = (0.5 * a + torch.randn(n, d, generator=g)).to(torch.bfloat16).float()
    a, bb = a.to(device), bb.to(device)
    cuda = a.is_cuda

    def one():
        if cuda and autocast_bf16:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return rank_share_step(a, bb, b, 1 / 0.07)
        return rank_share_step(a, bb, b, 1 / 0.07)

    for _ in range(warmup):
        one()
    if cuda:
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()  # float(loss) inside synchronises the device every step
    dt = time.perf_counter() - t0
    return dict(seconds=dt, ms_per_step=1e3 * dt / steps, samples_per_s=b * steps / dt,
                threads=torch.get_num_threads())
