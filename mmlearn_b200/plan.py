"""Host-side bookkeeping of the fused loss that needs no GPU: per-pair row counts across ranks,
label construction, and the coefficient table for the four (local_loss, gather_with_grad)
modes.  Kept free of CUDA so it is covered by CPU tests (single process and 2-rank gloo).

Notation: rank r holds b_r matched rows of a pair, N = sum_r b_r,
off_r = exclusive prefix sum.  With S the N x N logit matrix of the pair (rows: first modality,
columns: second), rank r owns row block [off_r, off_r + b_r).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import torch
import torch.distributed as dist


def dist_info() -> tuple[int, int]:
    """(world_size, rank) exactly as mmlearn/modules/losses/contrastive.py:86-87."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    return world, rank


def all_gather_ints(local: torch.Tensor, world: int) -> torch.Tensor:
    """All-gather a small int64 vector -> [world, len].  One collective replaces the
    reference's per-key shape gathers (contrastive.py:552-554) and the per-pair count gather
    (contrastive.py:197-206)."""
    if world == 1:
        return local.unsqueeze(0)
    out = torch.empty((world, local.numel()), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous())
    return out


@dataclass
class PairPlan:
    """Sizes and coefficients of one modality pair on one rank."""

    counts: Sequence[int]  # matched rows per rank (0 for a rank that skipped the pair)
    rank: int
    local_loss: bool
    gather_with_grad: bool
    weight: float = 1.0
    #: number of logit columns = globally matched pairs.  Equals sum(counts) unless example ids
    #: repeat across ranks, in which case the reference's labels are mis-aligned (SURVEY.md
    #: Appendix B) -- reproduced here bit for bit, and flagged by the caller.
    n_logits: Optional[int] = None

    @property
    def world(self) -> int:
        return len(self.counts)

    @property
    def n_total(self) -> int:
        return int(sum(self.counts))

    @property
    def b_local(self) -> int:
        return int(self.counts[self.rank])

    @property
    def offset(self) -> int:
        return int(sum(self.counts[: self.rank]))

    # ---- labels: contrastive.py:162-219 -------------------------------------------------
    def labels(self, device=None) -> torch.Tensor:
        """Ground-truth labels of this rank's logit rows (int64, bit-exact with the reference).

        ``arange(N)``, sliced to ``[off_r, off_{r+1})`` when ``world > 1 and local_loss``; the
        last rank's slice is open-ended exactly as at contrastive.py:210-212."""
        n = self.n_total if self.n_logits is None else int(self.n_logits)
        labels = torch.arange(n, device=device, dtype=torch.long)
        if self.world > 1 and self.local_loss:
            lo = self.offset
            hi = lo + self.b_local if self.rank + 1 < self.world else None
            labels = labels[lo:hi]
        return labels

    # ---- loss normalisation ---------------------------------------------------------------
    @property
    def global_loss(self) -> bool:
        """True when every rank reports the loss of the whole N x N problem
        (contrastive.py:334-337)."""
        return self.world > 1 and not self.local_loss

    @property
    def loss_rows(self) -> int:
        """Divisor of the cross-entropy mean: local rows, or all rows for the global loss."""
        return self.n_total if (self.world == 1 or self.global_loss) else self.b_local

    # ---- gradient coefficients --------------------------------------------------------------
    # The gradient a rank must return for its local rows is
    #   (T,T): d(sum_r loss_r)/dx   (F,T): W * d(global loss)/dx
    #   (T,F): d(loss_r)/dx with gathered copies constant      (F,F): d(global loss)/dx
    # In every case it has the form  s*w * sum_j (P_row[i,j]/den_row + P_col[i,j]/den_col(j)) y_j
    # minus the positive-label term, with P_row / P_col the row- / column-normalised softmax.
    def row_denominator(self) -> float:
        if self.world == 1:
            return 2.0 * self.n_total
        if self.local_loss:
            return 2.0 * max(self.b_local, 1)
        n = 2.0 * self.n_total
        return n / self.world if self.gather_with_grad else n

    def col_denominators(self) -> Optional[list]:
        """Per owning rank; ``None`` when the column term is absent ((T,F): keys carry no grad)."""
        if self.world == 1:
            return [2.0 * self.n_total]
        if self.local_loss:
            if not self.gather_with_grad:
                return None
            return [2.0 * max(int(c), 1) for c in self.counts]
        n = 2.0 * self.n_total
        return [n / self.world if self.gather_with_grad else n] * self.world

    def positive_coefficient(self) -> float:
        """Weight of the positive (label) column in G: 1/den_row + 1/den_col(own rank)."""
        c = 1.0 / self.row_denominator()
        cd = self.col_denominators()
        if cd is not None:
            c += 1.0 / cd[self.rank if self.world > 1 else 0]
        return c
