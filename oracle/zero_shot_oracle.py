"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the scoring step of the reference's zero-shot
classification task (mmlearn/tasks/zero_shot_classification.py:204-214).  Parity status: PINNED
by ``tests/golden/zeroshot_*.npz`` (``oracle/make_golden_zero_shot.py`` runs the unmodified
``evaluation_step``); only ``tests/`` import this module."""

from __future__ import annotations

import torch


def zero_shot_logits(query_embeddings: torch.Tensor, class_embeddings: torch.Tensor) -> torch.Tensor:
    """``logits`` as passed to ``metric_collection.update`` (:204-214)."""
    q = query_embeddings.float()
    q = q / q.norm(p=2, dim=-1, keepdim=True)  # :204
    cos = q @ class_embeddings.float().T  # _safe_matmul, :209-214
    if class_embeddings.shape[0] == 2:  # :207-212
        p = cos.softmax(dim=-1)
        return p[:, 1] - p[:, 0]
    return 100.0 * cos  # :214


def topk_hits(logits: torch.Tensor, targets: torch.Tensor, k: int) -> torch.Tensor:
    """Hit indicator of multiclass top-k accuracy: target among ``torch.topk(logits, k)``."""
    return (torch.topk(logits, k, dim=1).indices == targets[:, None]).any(1)
