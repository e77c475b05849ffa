"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32/fp64 + numpy) of the reference's
contrastive-loss path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module; the product
(``mmlearn_b200``) never does.

Parity status: PINNED.  ``oracle/make_golden.py`` runs the *unmodified* reference
(``/root/reference/mmlearn/modules/losses/contrastive.py`` via ``oracle/ref_loader.py``) in this
container, single-process and under 2-rank gloo, and stores inputs + outputs in
``tests/golden/*.npz``; ``tests/test_oracle.py`` checks this restatement against those files and
against the reference's own known-answer vectors for ``find_matching_indices``
(``/root/reference/tests/datasets/test_example.py:147-192``).  The two third-party pieces the
reference calls (torchmetrics 1.6.2 ``_safe_matmul``; torch 2.5.1 ``F.normalize`` /
``F.cross_entropy``) are restated from their published semantics.

Multi-rank semantics are restated *in one process*: "rank r" is an entry of a Python list, a
gather is a concatenation, and the gradient a rank ends up with is
  gather_with_grad=True : d(sum over ranks of loss_r)/d(local leaf)   (autograd all-gather's
                          backward is a reduce-scatter SUM, contrastive.py:506)
  gather_with_grad=False: d(loss_r)/d(local leaf); gathered copies are constants, except that
                          with local_loss=False the local shard is re-inserted (contrastive.py:491-492).
d(loss_r)/d(logit_scale) is always per rank (the scale is never gathered).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class LossPairSpec:
    """mmlearn/tasks/contrastive_pretraining.py:44-52."""

    modalities: tuple
    weight: float = 1.0


def find_matching_indices_np(first: np.ndarray, second: np.ndarray):
    """mmlearn/datasets/core/example.py:101-166 (broadcast compare + row-major ``where``)."""
    if not isinstance(first, np.ndarray) or not isinstance(second, np.ndarray):
        raise TypeError(
            f"Expected inputs to be tensors, but got {type(first)} and {type(second)}."
        )
    if not (first.ndim == 2 and first.shape[1] == 2):
        raise ValueError(
            "Expected argument `first_example_ids` to be a tensor of shape (N, 2), "
            f"but got shape {first.shape}."
        )
    if not (second.ndim == 2 and second.shape[1] == 2):
        raise ValueError(
            "Expected argument `second_example_ids` to be a tensor of shape (N, 2), "
            f"but got shape {second.shape}."
        )
    matches = np.all(first[:, None, :] == second[None, :, :], axis=-1)  # example.py:160-164
    ia, ib = np.nonzero(matches)  # row-major order == torch.where, example.py:165
    return ia.astype(np.int64), ib.astype(np.int64)


def find_matching_indices_sorted(first: np.ndarray, second: np.ndarray):
    """Same output as :func:`find_matching_indices_np` in O((n+m) log m) (for large-N checks)."""
    n, m = first.shape[0], second.shape[0]
    if n == 0 or m == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    order = np.lexsort((np.arange(m), second[:, 1], second[:, 0]))  # stable by (ds, ex, j)
    sk = second[order]
    skey = np.ascontiguousarray(sk).view([("a", sk.dtype), ("b", sk.dtype)]).ravel()
    fkey = np.ascontiguousarray(first).view([("a", first.dtype), ("b", first.dtype)]).ravel()
    lo = np.searchsorted(skey, fkey, side="left")
    hi = np.searchsorted(skey, fkey, side="right")
    cnt = hi - lo
    ia = np.repeat(np.arange(n, dtype=np.int64), cnt)
    starts = np.repeat(lo, cnt)
    within = np.arange(cnt.sum(), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    ib = order[starts + within].astype(np.int64)
    return ia, ib


def _emb_key(name: str) -> str:
    """mmlearn/datasets/core/modalities.py:51,66-82: ``Modality.embedding``."""
    return f"{name.lower()}_embedding"


@dataclass
class OracleResult:
    losses: list  # per rank: python float (fp64 of the torch scalar)
    grads: list  # per rank: dict emb_key -> torch.Tensor (same shape as the input)
    dscale: list  # per rank: float
    dscale_gross: list  # per rank: sum_ij |dL/dS_ij| |S_ij| / s  (conditioning scale of dscale)
    labels: list  # per rank: list (per pair) of np.int64 arrays (or None if the pair is skipped)
    matched: list  # per rank: list (per pair) of (idx_a, idx_b) local index arrays


def modality_alignment_oracle(all_embeddings: dict, logit_scale) -> torch.Tensor:
    """Literal restatement of ``_compute_modality_alignment_loss`` (contrastive.py:344-413),
    Python ``itertools.combinations`` loop and block-offset rule included."""
    import itertools

    keys = list(all_embeddings.keys())
    feats = torch.cat(list(all_embeddings.values()), dim=0)  # :374
    pos = []
    for idx, k in enumerate(all_embeddings):  # :376-386
        off = 0 if idx == 0 else all_embeddings[keys[idx - 1]].size(0)
        for i, j in itertools.combinations(range(all_embeddings[k].size(0)), 2):
            pos.append((i + off, j + off))
    logits = logit_scale * (feats @ feats.T)  # :387
    target = torch.eye(feats.size(0), dtype=logits.dtype)  # :389
    if pos:
        pi = torch.tensor(pos)
        target[pi[:, 0], pi[:, 1]] = 1  # :390
    bce = F.binary_cross_entropy_with_logits(logits, target, reduction="none")  # :392-394
    tp = target.bool()
    loss_pos = torch.where(tp, bce, torch.zeros_like(bce)).sum(dim=1)  # :401-409
    loss_neg = torch.where(~tp, bce, torch.zeros_like(bce)).sum(dim=1)
    num_pos = target.sum(dim=1)  # :410
    num_neg = logits.size(0) - num_pos  # :411
    return ((loss_pos / num_pos) + (loss_neg / num_neg)).mean()  # :413


def contrastive_oracle(
    emb_per_rank: Sequence[dict],
    ids_per_rank: Sequence[dict],
    logit_scale,
    pairs: Sequence,
    *,
    l2_normalize: bool = False,
    local_loss: bool = False,
    gather_with_grad: bool = False,
    dtype: torch.dtype = torch.float32,
    want_grads: bool = True,
    fast_match: bool = False,
    modality_alignment: bool = False,
) -> OracleResult:
    """Restates ``ContrastiveLoss.forward`` (contrastive.py:59-160) for W = len(emb_per_rank).

    ``logit_scale`` is a scalar (reference behaviour) or a sequence with one scale per pair (the
    new-API superset; additive over pairs).
    """
    W = len(emb_per_rank)
    match = find_matching_indices_sorted if fast_match else find_matching_indices_np
    P = len(pairs)
    s_in = torch.as_tensor(logit_scale, dtype=dtype).detach()
    per_pair_scale = s_in.ndim == 1
    # leaves
    leaves = [
        {k: v.detach().to(dtype).clone().requires_grad_(want_grads) for k, v in e.items()}
        for e in emb_per_rank
    ]
    scales = [s_in.clone().requires_grad_(want_grads) for _ in range(W)]
    # contrastive.py:89-90
    local = [
        {k: (F.normalize(v, p=2, dim=-1) if l2_normalize else v) for k, v in e.items()}
        for e in leaves
    ]
    ids_np = [{k: np.asarray(v, dtype=np.int64) for k, v in d.items()} for d in ids_per_rank]

    def gathered_for(rank: int) -> dict:
        """contrastive.py:431-499 with every rank holding every key."""
        if W == 1:
            return local[0]
        keys = sorted(set().union(*[set(e.keys()) for e in local]))
        out = {}
        for k in keys:
            parts = []
            for r in range(W):
                if k not in local[r]:
                    continue  # placeholder tensors are filtered out (contrastive.py:495)
                t = local[r][k]
                if gather_with_grad:
                    parts.append(t)
                elif r == rank and not local_loss:
                    parts.append(t)  # contrastive.py:491-492
                else:
                    parts.append(t.detach())
            out[k] = torch.cat(parts, dim=0)
        return out

    def gathered_ids() -> dict:
        keys = sorted(set().union(*[set(d.keys()) for d in ids_np]))
        return {
            k: np.concatenate([d[k] for d in ids_np if k in d], axis=0) for k in keys
        }

    all_ids = gathered_ids() if W > 1 else ids_np[0]

    losses_t, labels_out, matched_out = [], [], []
    gross = [0.0] * W
    for rank in range(W):
        all_emb = gathered_for(rank)
        emb = local[rank]
        rank_losses, rank_labels, rank_matched = [], [], []
        for p_idx, pair in enumerate(pairs):
            s = scales[rank][p_idx] if per_pair_scale else scales[rank]
            ma, mb = pair.modalities[0].lower(), pair.modalities[1].lower()
            ka, kb = _emb_key(ma), _emb_key(mb)
            skip = False
            fa = fb = None
            la = lb = None
            # contrastive.py:268-301
            if local_loss or W == 1:
                if not (ka in emb and kb in emb):
                    if W > 1:
                        skip = True
                    else:
                        rank_labels.append(None)
                        rank_matched.append(None)
                        continue
                if not skip:
                    la, lb = match(ids_np[rank][ma], ids_np[rank][mb])
                    if la.size == 0:
                        if W > 1:
                            skip = True
                        else:
                            rank_labels.append(None)
                            rank_matched.append(None)
                            continue
                if not skip:
                    fa = emb[ka][torch.from_numpy(la)]
                    fb = emb[kb][torch.from_numpy(lb)]
            # contrastive.py:303-340
            if W > 1:
                if not (ka in all_emb and kb in all_emb):
                    rank_labels.append(None)
                    rank_matched.append(None)
                    continue
                ga, gb = match(all_ids[ma], all_ids[mb])
                if ga.size == 0:
                    rank_labels.append(None)
                    rank_matched.append(None)
                    continue
                afa = all_emb[ka][torch.from_numpy(ga)]
                afb = all_emb[kb][torch.from_numpy(gb)]
                if local_loss:
                    if fa is None:
                        fa, fb = afa, afb
                    logits_a = s * (fa @ afb.T)
                    logits_b = s * (fb @ afa.T)
                else:
                    logits_a = s * (afa @ afb.T)
                    logits_b = logits_a.T
            else:
                logits_a = s * (fa @ fb.T)
                logits_b = s * (fb @ fa.T)
            # contrastive.py:162-219 (labels; the cache only changes *when* they are rebuilt)
            num_logits = logits_a.shape[-1]
            labels = np.arange(num_logits, dtype=np.int64)
            if W > 1 and local_loss:
                sizes = []
                for r in range(W):
                    if ka in local[r] and kb in local[r]:
                        ra, _ = match(ids_np[r][ma], ids_np[r][mb])
                        sizes.append(ra.size)
                    else:
                        sizes.append(0)
                cum = np.concatenate([[0], np.cumsum(sizes)])
                labels = labels[cum[rank]: (cum[rank + 1] if rank + 1 < W else None)]
            rank_labels.append(labels)
            rank_matched.append((la, lb) if la is not None else None)
            if labels.size != 0:
                lt = torch.from_numpy(labels)
                rank_losses.append(
                    (F.cross_entropy(logits_a, lt) + F.cross_entropy(logits_b, lt))
                    / 2 * pair.weight
                )  # contrastive.py:134-144
                with torch.no_grad():
                    for lg in (logits_a, logits_b):
                        g = torch.softmax(lg, dim=1)
                        g[torch.arange(lt.numel()), lt] -= 1.0
                        gross[rank] += float(
                            (g.abs() * lg.abs()).sum() * abs(pair.weight) / (2 * lt.numel())
                            / abs(float(s)))
        if modality_alignment:  # contrastive.py:146-149
            rank_losses.append(modality_alignment_oracle(all_emb, scales[rank]))
        if not rank_losses:
            rank_losses.append(torch.zeros((), dtype=dtype))  # contrastive.py:151-158
        losses_t.append(torch.stack(rank_losses).sum())
        labels_out.append(rank_labels)
        matched_out.append(rank_matched)

    grads = [dict() for _ in range(W)]
    dscale = [0.0] * W
    if want_grads:
        total = torch.stack(losses_t).sum()
        if total.requires_grad:
            total.backward()
        for r in range(W):
            for k, v in leaves[r].items():
                grads[r][k] = v.grad.clone() if v.grad is not None else torch.zeros_like(v)
            g = scales[r].grad
            dscale[r] = (
                g.clone().double().numpy() if g is not None else np.zeros(s_in.shape)
            )
    return OracleResult(
        losses=[float(l.detach().double()) for l in losses_t],
        grads=grads,
        dscale=dscale,
        dscale_gross=gross,
        labels=labels_out,
        matched=matched_out,
    )


def synthetic_inputs(
    n: int,
    d: int,
    modalities: Sequence[str] = ("rgb", "text"),
    world: int = 1,
    seed: int = 1234,
    correlated: bool = True,
    bf16_round: bool = True,
):
    """SURVEY.md section 8d synthetic inputs: per-modality N(0,1) rows (optionally correlated with
    the first modality), rounded to bf16-representable fp32, ids ``(0, global_row)``; rank r gets
    rows ``[r*b, (r+1)*b)``."""
    assert n % world == 0
    b = n // world
    base = None
    embs = {}
    for mi, m in enumerate(modalities):
        g = torch.Generator().manual_seed(seed + mi)
        x = torch.randn(n, d, generator=g, dtype=torch.float32)
        if correlated and base is not None:
            x = 0.5 * base + x
        if base is None:
            base = x
        if bf16_round:
            x = x.to(torch.bfloat16).to(torch.float32)
        embs[m] = x
    emb_per_rank, ids_per_rank = [], []
    for r in range(world):
        emb_per_rank.append({_emb_key(m): embs[m][r * b:(r + 1) * b].clone() for m in modalities})
        ids = np.stack([np.zeros(b, np.int64), np.arange(r * b, (r + 1) * b, dtype=np.int64)], 1)
        ids_per_rank.append({m: ids.copy() for m in modalities})
    return emb_per_rank, ids_per_rank
