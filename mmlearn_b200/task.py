"""Caller glue of ``ContrastivePretraining`` around the loss.

The reference task does, every step (mmlearn/tasks/contrastive_pretraining.py):

* ``encode(..., normalize=True)`` -> ``F.normalize`` per modality (:428-429, :447), and the loss
  normalises again when ``l2_normalize=True`` (contrastive.py:89-90);
* ``self.log_logit_scale.clamp_(0, log(max_logit_scale))`` under ``no_grad`` (:488-489);
* ``self.loss_fn(outputs, batch["example_ids"], self.log_logit_scale.exp(), pairs)`` (:607-612).

``ContrastiveLossHead`` owns the ``log_logit_scale`` parameter exactly like the task
(:327-337: init ``log(1/0.07)``, learnable or buffer, one scalar -- or, as the new-API superset,
one per pair) and takes the encoders' RAW outputs: the normalisation forward and backward live in
the pack / unpack kernels of the fused loss (the rows are never normalised in HBM), and clamp + exp
are one launch (``mmcl_logit_scale``).  Because normalisation is idempotent, feeding it already
normalised outputs gives the same numbers, so it can replace ``_compute_loss`` of an unchanged task.
"""

from __future__ import annotations

import itertools
import math
from typing import Optional, Sequence

import torch
from torch import nn

from . import _lib
from ._lib import check
from .loss import ContrastiveLoss, _stream, get_handle
from .modalities import LossPairSpec


class _ClampExp(torch.autograd.Function):
    """s = exp(clamp_(log_s)); the clamp is applied to the parameter in place without entering
    the graph (contrastive_pretraining.py:488-489), so d s / d log_s = s."""

    @staticmethod
    def forward(ctx, log_s: torch.Tensor, max_scale: float):
        if not log_s.is_cuda or log_s.dtype != torch.float32 or not log_s.is_contiguous():
            raise ValueError("log_logit_scale must be a contiguous fp32 CUDA tensor")
        s = torch.empty_like(log_s)
        check(_lib.load().mmcl_logit_scale(get_handle(log_s.device), log_s.data_ptr(),
                                           log_s.numel(), float(max_scale), s.data_ptr(),
                                           _stream(log_s.device)), "mmcl_logit_scale")
        ctx.save_for_backward(s)
        return s

    @staticmethod
    def backward(ctx, g):
        (s,) = ctx.saved_tensors
        return g * s, None


class ContrastiveLossHead(nn.Module):
    """``log_logit_scale`` + loss call of the reference task, fused.

    Parameters follow ``ContrastivePretraining.__init__`` (contrastive_pretraining.py:150-165):
    ``init_logit_scale=1/0.07``, ``max_logit_scale=100``, ``learnable_logit_scale=True``,
    ``modality_loss_pairs`` (default: all combinations of the modalities seen in the first
    call).  ``per_pair_scale=True`` keeps one scale per pair instead of the reference's single
    scalar.  ``loss`` defaults to the fused loss with ``l2_normalize=True``."""

    def __init__(self, loss: Optional[nn.Module] = None, init_logit_scale: float = 1 / 0.07,
                 max_logit_scale: float = 100.0, learnable_logit_scale: bool = True,
                 modality_loss_pairs: Optional[Sequence[LossPairSpec]] = None,
                 per_pair_scale: bool = False):
        super().__init__()
        self.loss_fn = loss if loss is not None else ContrastiveLoss(
            l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=True)
        self.max_logit_scale = max_logit_scale
        self.learnable_logit_scale = learnable_logit_scale
        self.modality_loss_pairs = list(modality_loss_pairs) if modality_loss_pairs else None
        n = len(self.modality_loss_pairs) if (per_pair_scale and self.modality_loss_pairs) else 0
        if per_pair_scale and not n:
            raise ValueError("per_pair_scale needs explicit modality_loss_pairs")
        log_s = torch.ones([n] if n else []) * math.log(init_logit_scale)
        if learnable_logit_scale:
            self.log_logit_scale = nn.Parameter(log_s, requires_grad=True)
        else:
            self.register_buffer("log_logit_scale", log_s)

    def logit_scale(self) -> torch.Tensor:
        """Training: clamp the parameter in place and return exp of it (one kernel), as
        ``training_step`` does (contrastive_pretraining.py:488-489).  Evaluation: the reference
        only exponentiates (``_shared_eval_step`` -> ``_compute_loss``, :607-611) -- the parameter
        is left untouched."""
        if not self.training:
            return self.log_logit_scale.exp()
        return _ClampExp.apply(self.log_logit_scale, self.max_logit_scale)

    def forward(self, outputs: dict, example_ids: dict) -> torch.Tensor:
        """``outputs``: ``{"<modality>_embedding": [b, d]}`` raw (or normalised) encoder outputs."""
        pairs = self.modality_loss_pairs
        if pairs is None:  # contrastive_pretraining.py:339-344
            names = [k[:-len("_embedding")] for k in outputs]
            pairs = [LossPairSpec(modalities=(a, b)) for a, b in itertools.combinations(names, 2)]
        return self.loss_fn(outputs, example_ids, self.logit_scale(), pairs)
