"""The ``modality_alignment`` term of the reference loss
(mmlearn/modules/losses/contrastive.py:344-413), kept so that the constructor surface of the
drop-in is complete.

SURVEY.md section 8 (row a11) places this term OUTSIDE the fused hot path: the reference builds its
positive list with an O(N^2) Python loop and a dense [sum N, sum N] all-vs-all matrix, so it is
only usable at toy sizes.  This module therefore is a plain dense restatement on the caller's
CUDA tensors (torch ops; no hand-written kernel, no CPU fallback), differentiable through
autograd, with the reference's exact target construction -- including its quirks:

* only the upper triangle (i < j) of a modality's block, plus the whole diagonal, is positive;
* the block of the k-th modality (k >= 1) starts at ``size(modality k-1)``, not at the cumulative
  offset (contrastive.py:375-386) -- identical for two modalities, different for three.

The positive list is built as a mask (``triu``), not as Python tuples.
"""

from __future__ import annotations

import torch
import torch.distributed as dist
import torch.nn.functional as F

from .plan import dist_info


def alignment_targets(sizes: list, device=None) -> torch.Tensor:
    """Target matrix of contrastive.py:375-391 for modalities with ``sizes`` rows (fp32 0/1)."""
    n = int(sum(sizes))
    target = torch.eye(n, device=device)
    for k, nk in enumerate(sizes):
        off = 0 if k == 0 else int(sizes[k - 1])
        if nk > 1:
            blk = target[off:off + nk, off:off + nk]
            blk.copy_(torch.maximum(blk, torch.triu(torch.ones(nk, nk, device=device), 1)))
    return target


def _gather_rows(t: torch.Tensor, with_grad: bool, reinsert_local: bool) -> torch.Tensor:
    """``_gather_dicts`` for one tensor (contrastive.py:431-499, 513-578): rank-ordered
    concatenation of every rank's rows; uneven row counts are padded to the maximum and sliced."""
    world, rank = dist_info()
    sizes = torch.zeros(world, dtype=torch.long, device=t.device)
    sizes[rank] = t.shape[0]
    dist.all_reduce(sizes)
    sizes = sizes.tolist()
    mx = max(sizes)
    padded = t if t.shape[0] == mx else F.pad(t, (0, 0, 0, mx - t.shape[0]))
    padded = padded.contiguous()
    if with_grad:
        import torch.distributed.nn.functional as dist_fn

        parts = list(dist_fn.all_gather(padded))
    else:
        parts = [torch.empty_like(padded) for _ in range(world)]
        dist.all_gather(parts, padded.detach())
    parts = [p[:n] for p, n in zip(parts, sizes)]
    if reinsert_local:
        parts[rank] = t
    return torch.cat(parts, 0)


def modality_alignment_loss(embeddings: dict, logit_scale: torch.Tensor, l2_normalize: bool,
                            local_loss: bool, gather_with_grad: bool) -> torch.Tensor:
    """``_compute_modality_alignment_loss`` on (gathered) embeddings; fp32 arithmetic."""
    world, _ = dist_info()
    if logit_scale.numel() != 1:
        raise ValueError("modality_alignment needs a scalar logit_scale (as in the reference)")
    embs = {k: (F.normalize(v.float(), p=2, dim=-1) if l2_normalize else v.float())
            for k, v in embeddings.items()}
    if world > 1:
        # key order of _gather_dicts: sorted union (contrastive.py:464-467); every rank is
        # expected to hold every key (the placeholder path is part of the fused term only)
        embs = {k: _gather_rows(embs[k], gather_with_grad,
                                reinsert_local=not gather_with_grad and not local_loss)
                for k in sorted(embs)}
    feats = torch.cat(list(embs.values()), 0)
    target = alignment_targets([v.shape[0] for v in embs.values()], feats.device)
    logits = logit_scale.float().reshape(()) * (feats @ feats.T)
    bce = F.binary_cross_entropy_with_logits(logits, target, reduction="none")
    pos = target > 0
    loss_pos = (bce * pos).sum(1)
    loss_neg = (bce * ~pos).sum(1)
    num_pos = target.sum(1)
    num_neg = logits.shape[0] - num_pos
    return ((loss_pos / num_pos) + (loss_neg / num_neg)).mean()
