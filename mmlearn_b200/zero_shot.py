"""B200-native scoring step of ``ZeroShotClassification``
(mmlearn/tasks/zero_shot_classification.py:179-218, ``evaluation_step``).

The reference normalises the query embeddings, multiplies them with the (normalised) class
embeddings and hands ``100 * cos`` (or, for two classes, ``softmax(cos)[:, 1] - softmax(cos)[:, 0]``)
together with the targets to a torchmetrics collection.  Here both products run on the tcgen05
similarity mainloop of the loss:

* :meth:`ZeroShotScores.logits` -- the scores themselves (``mmcl_sim_logits_fwd``: the forward
  kernel with a storing epilogue), for callers that feed a metric collection;
* :meth:`ZeroShotScores.target_ranks` -- per query, how many classes score above the target class
  (``mmcl_sim_rank_fwd``: the counting epilogue, nothing stored); ``rank < k`` is top-k accuracy's
  hit indicator for every ``k`` at once (the reference builds one ``Accuracy`` metric per ``k``,
  zero_shot_classification.py:241-250).

Arithmetic: rows are rounded to bf16 and L2-normalised by fp32 factors in the epilogue, products
accumulate in fp32 (same contract as the loss).  Classes whose score ties with
the target's (within ``tie_margin``) count in favour of the target.
"""

from __future__ import annotations

from typing import Optional

import torch

from . import _lib
from ._lib import check, ptr


class ZeroShotScores:
    """Scores of query embeddings against a fixed set of class embeddings.

    Parameters
    ----------
    class_embeddings : torch.Tensor
        ``[num_classes, d]`` CUDA tensor: what ``on_evaluation_epoch_start`` stores per dataset
        (zero_shot_classification.py:131-177), i.e. already normalised.  Used as they are, like the
        reference does; pass ``normalize_classes=True`` for raw class embeddings.
    scale : float, default=100.0
        Multiplier of the cosine similarities (zero_shot_classification.py:214).
    tie_margin : float, default=1e-6
        See the module docstring.
    """

    def __init__(self, class_embeddings: torch.Tensor, scale: float = 100.0,
                 tie_margin: float = 1e-6, normalize_classes: bool = False):
        if not isinstance(class_embeddings, torch.Tensor) or not class_embeddings.is_cuda:
            raise ValueError("mmlearn_b200.ZeroShotScores needs CUDA tensors (no CPU fallback)")
        if class_embeddings.ndim != 2 or class_embeddings.shape[1] % 8 != 0 or \
                class_embeddings.shape[1] > 2048:
            raise ValueError("class_embeddings must be [num_classes, d] with d a multiple of 8 "
                             f"(<= 2048), got {tuple(class_embeddings.shape)}")
        from .loss import get_handle

        self.dev = class_embeddings.device
        self.h = get_handle(self.dev)
        self.scale = float(scale)
        self.tie_margin = float(tie_margin)
        self.num_classes, self.d = class_embeddings.shape
        self.k, self.kinv = self._pack(class_embeddings)
        if not normalize_classes:
            self.kinv = None  # factor 1: the stored embeddings are used as they are

    def _pack(self, x: torch.Tensor):
        from .loss import _stream

        lib = _lib.load()
        x = x.detach().to(self.dev).float().contiguous()
        n = x.shape[0]
        rows = torch.empty(n, self.d, dtype=torch.bfloat16, device=self.dev)
        inv = torch.empty(n, dtype=torch.float32, device=self.dev)
        check(lib.mmcl_pack_rows(self.h, ptr(x), 0, n, self.d, None, n, 1, ptr(rows), ptr(inv),
                                 _stream(self.dev)), "mmcl_pack_rows")
        return rows, inv

    def _check_queries(self, q: torch.Tensor) -> None:
        if not isinstance(q, torch.Tensor) or not q.is_cuda:
            raise ValueError("mmlearn_b200.ZeroShotScores needs CUDA tensors (no CPU fallback)")
        if q.ndim != 2 or q.shape[1] != self.d:
            raise ValueError(f"query embeddings must be [batch, {self.d}], got {tuple(q.shape)}")

    def logits(self, query_embeddings: torch.Tensor) -> torch.Tensor:
        """What the reference passes to ``metric_collection.update``: ``[batch, num_classes]``
        scores ``scale * cos``, or for two classes the ``[batch]`` vector
        ``softmax(cos)[:, 1] - softmax(cos)[:, 0]`` (zero_shot_classification.py:208-214)."""
        from .loss import _stream

        self._check_queries(query_embeddings)
        lib = _lib.load()
        n = query_embeddings.shape[0]
        binary = self.num_classes == 2
        out = torch.empty(n, self.num_classes, dtype=torch.float32, device=self.dev)
        if n == 0:
            return out[:, 0] if binary else out
        q, qinv = self._pack(query_embeddings)
        check(lib.mmcl_sim_logits_fwd(self.h, ptr(q), n, ptr(self.k), self.num_classes, self.d,
                                      1.0 if binary else self.scale, ptr(qinv), ptr(self.kinv),
                                      ptr(out), self.num_classes, _stream(self.dev)),
              "mmcl_sim_logits_fwd")
        if binary:
            p = out.softmax(dim=-1)
            return p[:, 1] - p[:, 0]
        return out

    def target_ranks(self, query_embeddings: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
        """int64 ``[batch]``: the number of classes that score above each query's target class."""
        from .loss import _stream

        self._check_queries(query_embeddings)
        lib = _lib.load()
        n = query_embeddings.shape[0]
        out = torch.empty(n, dtype=torch.long, device=self.dev)
        if n == 0:
            return out
        tg = targets.detach().to(device=self.dev, dtype=torch.long).contiguous()
        if tg.shape != (n,):
            raise ValueError(f"targets must be [{n}], got {tuple(tg.shape)}")
        q, qinv = self._pack(query_embeddings)
        st = _stream(self.dev)
        thr = torch.empty(n, dtype=torch.float32, device=self.dev)
        check(lib.mmcl_indexed_similarity(self.h, ptr(q), n, ptr(self.k), self.num_classes, self.d,
                                          ptr(tg), ptr(qinv), ptr(self.kinv), self.tie_margin,
                                          ptr(thr), st), "mmcl_indexed_similarity")
        ws_bytes = lib.mmcl_sim_lse_workspace_bytes(self.h, n, self.num_classes, self.d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=self.dev)
        check(lib.mmcl_sim_rank_fwd(self.h, ptr(q), n, ptr(self.k), self.num_classes, self.d,
                                    ptr(qinv), ptr(self.kinv), ptr(thr), ptr(tg), ptr(out), ptr(ws),
                                    ws_bytes, st), "mmcl_sim_rank_fwd")
        return out

    def topk_correct(self, query_embeddings: torch.Tensor, targets: torch.Tensor,
                     k: int = 1, ranks: Optional[torch.Tensor] = None) -> torch.Tensor:
        """bool ``[batch]``: target among the ``k`` best-scoring classes (micro top-k accuracy is
        its mean)."""
        r = self.target_ranks(query_embeddings, targets) if ranks is None else ranks
        return r < int(k)
