"""B200-native ``RetrievalRecallAtK`` (mmlearn/modules/metrics/retrieval_recall.py:20-289).

Same constructor, ``update(x, y, indexes)`` / ``compute()`` / ``reset()`` contract as the reference
metric, but nothing is moved to the CPU and no similarity matrix, ``topk`` or sort exists: the
query rows go through the tcgen05 similarity mainloop of the loss (``mmcl_sim_rank_fwd``) whose
epilogue counts, per query, the database rows that beat the positive.  ``rank < k`` is the hit
indicator the reference derives from ``topk`` (``_recall_at_k``), so one pass serves every ``k``
(``ranks()``; the reference's zero-shot retrieval task instantiates one metric per k and repeats
the whole matmul + topk for each, zero_shot_retrieval.py:63-70).

Arithmetic: rows are rounded to bf16 and L2-normalised by fp32 factors in the epilogue, products
accumulate in fp32.  Ties (another row within ``tie_margin`` of the positive's cosine) count in
favour of the positive; ``torch.topk`` resolves exact ties by position.

Distributed: every rank keeps its own queries; ``compute`` all-gathers the database rows (NCCL),
ranks the local queries against all of them and all-reduces the hit count -- the reference instead
gathers everything onto every rank in ``update`` (:139-160) and repeats the full computation on
each.  The global row order it defines (step-major, then rank) is reproduced by ``global_layout``.
"""

from __future__ import annotations

from typing import Any, Callable, Optional, Sequence, Union

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from ._lib import check, ptr
from .plan import dist_info


def global_layout(sizes: Sequence[Sequence[int]], rank: int):
    """Row bookkeeping of the reference's distributed ``update`` (retrieval_recall.py:137-160).

    ``sizes[t][r]`` = rows rank r contributed at update step t.  Returns
    ``(offsets, order)``: ``offsets[t]`` is what must be added to rank ``rank``'s local
    ``indexes`` of step t (``num_samples`` so far + rows of lower ranks in this step), and
    ``order`` lists ``(rank, step)`` blocks in global row order (step-major, then rank)."""
    offsets, order, seen = [], [], 0
    for t, row in enumerate(sizes):
        offsets.append(seen + int(sum(row[:rank])))
        order.extend((r, t) for r in range(len(row)))
        seen += int(sum(row))
    return offsets, order


class RetrievalRecallAtK(nn.Module):
    """Retrieval Recall@K (fused B200 implementation); parameters as in the reference
    (retrieval_recall.py:29-53)."""

    is_differentiable: bool = False
    higher_is_better: bool = True
    full_state_update: bool = False

    def __init__(
        self,
        top_k: int,
        reduction: Optional[str] = "sum",
        aggregation: Union[str, Callable[[torch.Tensor, int], torch.Tensor]] = "mean",
        tie_margin: float = 1e-6,
        **kwargs: Any,
    ) -> None:
        super().__init__()
        if top_k is not None and not (isinstance(top_k, int) and top_k > 0):
            raise ValueError("`top_k` has to be a positive integer or None")
        self.top_k = top_k
        allowed_reduction = ("sum", "mean", "none", None)
        if reduction not in allowed_reduction:
            raise ValueError(
                f"Expected argument `reduction` to be one of {allowed_reduction} but got {reduction}"
            )
        self.reduction = reduction
        if not (aggregation in ("mean", "median", "min", "max") or callable(aggregation)):
            raise ValueError(
                "Argument `aggregation` must be one of `mean`, `median`, `min`, `max` or a custom callable function"
                f"which takes tensor of values, but got {aggregation}."
            )
        self.aggregation = aggregation
        self.tie_margin = float(tie_margin)
        self.reset()

    # ------------------------------------------------------------------------------------
    def reset(self) -> None:
        self.x: list = []
        self.y: list = []
        self.indexes: list = []
        self._ranks: Optional[torch.Tensor] = None

    def update(self, x: torch.Tensor, y: torch.Tensor, indexes: torch.Tensor) -> None:
        """Accumulate a batch (checks of ``_update_batch_inputs``, retrieval_recall.py:292-333)."""
        if indexes is None:
            raise ValueError("Argument `indexes` cannot be None")
        if x.shape != y.shape:
            raise RuntimeError(
                "Predictions and targets are expected to have the same shape, "
                f"but got {x.shape} and {y.shape}.")
        if x.ndim != 2:
            raise ValueError(
                "Expected input to retrieval recall to be 2D tensors of shape `[N,D]`, "
                "where `N` is the number of samples and `D` is the number of dimensions, "
                f"but got tensor of shape {x.shape}"
            )
        if not indexes.numel() or not indexes.size():
            raise ValueError("`indexes`, `x` and `y` must be non-empty and non-scalar tensors")
        if not x.is_cuda:
            raise ValueError("mmlearn_b200.RetrievalRecallAtK needs CUDA tensors (no CPU fallback)")
        if x.shape[1] % 8 != 0 or x.shape[1] > 2048:
            raise ValueError("embedding dimension must be a multiple of 8 (<= 2048)")
        self.x.append(x.detach().float())
        self.y.append(y.detach().float())
        self.indexes.append(indexes.detach().long().to(x.device))
        self._ranks = None

    # ------------------------------------------------------------------------------------
    def ranks(self) -> torch.Tensor:
        """int64 [n_local_queries]: how many database rows beat each query's positive."""
        if self._ranks is not None:
            return self._ranks
        if not self.x:
            raise ValueError("No samples to concatenate")
        from .loss import _stream, get_handle

        lib = _lib.load()
        world, rank = dist_info()
        dev = self.x[0].device
        h = get_handle(dev)
        st = _stream(dev)
        x = torch.cat(self.x)
        y = torch.cat(self.y)
        n, d = x.shape
        if world == 1:
            off = 0
            idx_parts = []
            for xb, ib in zip(self.x, self.indexes):
                idx_parts.append(ib + off)  # retrieval_recall.py:137
                off += xb.shape[0]
            idx = torch.cat(idx_parts)
            y_all = y
        else:
            # sizes[t][r] of every update step on every rank (all ranks update in lock step)
            mine = torch.tensor([xb.shape[0] for xb in self.x], device=dev, dtype=torch.long)
            sizes = torch.empty(world, mine.numel(), dtype=torch.long, device=dev)
            dist.all_gather_into_tensor(sizes.view(-1), mine)
            sizes_l = sizes.t().tolist()
            offsets, order = global_layout(sizes_l, rank)
            idx = torch.cat([ib + o for ib, o in zip(self.indexes, offsets)])
            per_rank = sizes.sum(1).tolist()
            mx = max(per_rank)
            pad = y if y.shape[0] == mx else torch.nn.functional.pad(y, (0, 0, 0, mx - y.shape[0]))
            gathered = torch.empty(world, mx, d, device=dev, dtype=y.dtype)
            dist.all_gather_into_tensor(gathered.view(-1), pad.contiguous().view(-1))
            starts = [[sum(sizes_l[tt][r] for tt in range(t)) for t in range(len(sizes_l))]
                      for r in range(world)]
            y_all = torch.cat([gathered[r, starts[r][t]:starts[r][t] + sizes_l[t][r]]
                               for r, t in order])
        m = y_all.shape[0]
        q = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
        k = torch.empty(m, d, dtype=torch.bfloat16, device=dev)
        qinv = torch.empty(n, dtype=torch.float32, device=dev)
        kinv = torch.empty(m, dtype=torch.float32, device=dev)
        xc, yc = x.contiguous(), y_all.contiguous()
        check(lib.mmcl_pack_rows(h, ptr(xc), 0, n, d, None, n, 1, ptr(q), ptr(qinv), st),
              "mmcl_pack_rows")
        check(lib.mmcl_pack_rows(h, ptr(yc), 0, m, d, None, m, 1, ptr(k), ptr(kinv), st),
              "mmcl_pack_rows")
        thr = torch.empty(n, dtype=torch.float32, device=dev)
        idx = idx.contiguous()
        check(lib.mmcl_indexed_similarity(h, ptr(q), n, ptr(k), m, d, ptr(idx), ptr(qinv),
                                          ptr(kinv), self.tie_margin, ptr(thr), st),
              "mmcl_indexed_similarity")
        ws_bytes = lib.mmcl_sim_lse_workspace_bytes(h, n, m, d)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out = torch.empty(n, dtype=torch.long, device=dev)
        check(lib.mmcl_sim_rank_fwd(h, ptr(q), n, ptr(k), m, d, ptr(qinv), ptr(kinv), ptr(thr),
                                    ptr(idx), ptr(out), ptr(ws), ws_bytes, st),
              "mmcl_sim_rank_fwd")
        self._ranks = out
        return out

    def hits(self, top_k: Optional[int] = None) -> torch.Tensor:
        """bool [n_local_queries]: positive among the top-k (``_recall_at_k(...) > 0``)."""
        return self.ranks() < int(self.top_k if top_k is None else top_k)

    def compute(self, top_k: Optional[int] = None) -> torch.Tensor:
        """The metric (retrieval_recall.py:172-223): aggregation over all queries of all ranks."""
        if self.reduction in ("sum", "mean"):
            # the reference reduces the score matrix to one value per query before topk(dim=1)
            # (:259-260) and fails there; only reduction="none" is usable (as its task does)
            raise NotImplementedError(
                "reduction='sum'/'mean' collapses the score matrix before top-k in the reference "
                "(and raises there); use reduction='none'")
        vals = self.hits(top_k).float()
        world, _ = dist_info()
        if world > 1:
            if self.aggregation == "mean":
                s = torch.stack([vals.sum(), torch.tensor(float(vals.numel()), device=vals.device)])
                dist.all_reduce(s)
                return s[0] / s[1]
            sizes = torch.zeros(world, dtype=torch.long, device=vals.device)
            sizes[dist.get_rank()] = vals.numel()
            dist.all_reduce(sizes)
            mx = int(sizes.max())
            buf = torch.empty(world, mx, device=vals.device)
            dist.all_gather_into_tensor(buf.view(-1),
                                        torch.nn.functional.pad(vals, (0, mx - vals.numel())))
            vals = torch.cat([buf[r, :int(sizes[r])] for r in range(world)])
        if self.aggregation == "mean":
            return vals.mean()
        if self.aggregation == "median":
            return vals.median()
        if self.aggregation == "min":
            return vals.min()
        if self.aggregation == "max":
            return vals.max()
        return self.aggregation(vals, dim=None)

    def forward(self, *args: Any, **kwargs: Any) -> Any:
        """Forward method is not supported (as in the reference, retrieval_recall.py:225-236)."""
        raise NotImplementedError("RetrievalRecallAtK metric does not support forward method")


try:  # pragma: no cover - hydra-zen / mmlearn are not installable in the build image
    from mmlearn.conf import external_store  # type: ignore

    # the reference registers its metric in group "modules/metrics" (retrieval_recall.py:20)
    external_store(RetrievalRecallAtK, name="FusedRetrievalRecallAtK", group="modules/metrics",
                   provider="mmlearn_b200")
except Exception:  # noqa: BLE001
    pass
