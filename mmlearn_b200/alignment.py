"""The ``modality_alignment`` term of the reference loss
(mmlearn/modules/losses/contrastive.py:344-413) on the tensor-core mainloops of the fused loss.

The reference concatenates the (gathered) embeddings of all modalities, forms the all-vs-all
``[T, T]`` logit matrix, builds a dense 0/1 target from an O(T^2) Python list of index pairs and
applies binary cross-entropy with logits.  Its targets have a simple structure -- including its
quirks:

* only the upper triangle (i < j) of a modality's block, plus the whole diagonal, is positive;
* the block of the k-th modality (k >= 1) starts at ``size(modality k-1)``, not at the cumulative
  offset (contrastive.py:375-386) -- identical for two modalities, different for three;

so the positives of row ``i`` are one contiguous column band ``[i, band_end[i])``
(:func:`band_ends`).  The fused path evaluates that band from tile coordinates:

* forward: ``mmcl_sim_bce_fwd`` (the similarity mainloop of the loss with a softplus / sigmoid
  epilogue) gives, per local row, the five sums the loss and d(loss)/d(scale) need;
* backward: ``mmcl_sim_bce_bwd`` (the recompute-backward kernel with the BCE form of G) gives the
  complete gradient of every local row -- each rank needs only the gathered rows, never a
  ``[T, d]`` gradient exchange;

nothing of size ``[T, T]`` exists.  Rows are rounded to bf16 and normalised by fp32 factors in the
epilogues (the loss's precision contract); the embedding width is zero-padded to what the CTA-pair
backward kernel needs (a multiple of 128).

Distributed semantics are those of ``_gather_dicts`` (contrastive.py:431-499): keys are the sorted
union over ranks, rows are concatenated in rank order, uneven row counts are allowed; the gradient
reaching the local rows is ``world x`` the single-process one with ``gather_with_grad``, the
single-process one when the local shard is re-inserted (``not gather_with_grad and not
local_loss``), and none otherwise.  A rank that lacks a modality contributes zero rows.
"""

from __future__ import annotations

from typing import Optional, Sequence

import torch
import torch.distributed as dist

from . import _lib
from ._lib import check, counts_array, ptr
from .plan import dist_info

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def alignment_targets(sizes: Sequence[int], device=None) -> torch.Tensor:
    """Dense target matrix of contrastive.py:375-391 for modalities with ``sizes`` rows (fp32
    0/1).  Test / documentation helper: the fused path never builds it."""
    n = int(sum(sizes))
    target = torch.eye(n, device=device)
    for k, nk in enumerate(sizes):
        off = 0 if k == 0 else int(sizes[k - 1])
        if nk > 1:
            blk = target[off:off + nk, off:off + nk]
            blk.copy_(torch.maximum(blk, torch.triu(torch.ones(nk, nk, device=device), 1)))
    return target


def band_ends(sizes: Sequence[int]) -> list:
    """``band_end[i]`` such that the positives of row ``i`` are exactly the columns
    ``[i, band_end[i])``: the diagonal, and ``i < j < off_k + n_k`` for every modality block
    ``[off_k, off_k + n_k)`` that contains ``i`` (``off_k`` as the reference computes it)."""
    total = int(sum(sizes))
    end = [i + 1 for i in range(total)]
    for k, nk in enumerate(sizes):
        off = 0 if k == 0 else int(sizes[k - 1])
        hi = min(off + int(nk), total)
        for i in range(off, hi):
            end[i] = max(end[i], hi)
    return end


def _padded_width(d: int) -> int:
    """Smallest width >= d the CTA-pair backward kernel supports."""
    dp = (d + 127) // 128 * 128
    while dp % (128 * ((dp + 511) // 512)) != 0:
        dp += 128
    return dp


class _FusedAlignmentFn(torch.autograd.Function):
    """loss = f(logit_scale, *embeddings); ``meta`` carries the non-tensor arguments."""

    @staticmethod
    def forward(ctx, meta, logit_scale, *embs):
        from .loss import _stream, get_handle

        lib = _lib.load()
        world, rank = dist_info()
        keys, all_keys, l2_normalize, grad_factor = meta
        dev = logit_scale.device if logit_scale.is_cuda else embs[0].device
        h = get_handle(dev)
        st = _stream(dev)
        by_key = dict(zip(keys, embs))
        d = embs[0].shape[-1]
        dp = _padded_width(d)
        # rows of every (modality, rank): one small exchange
        mine = torch.tensor([by_key[k].shape[0] if k in by_key else 0 for k in all_keys],
                            dtype=torch.long, device=dev)
        if world > 1:
            counts_t = torch.empty(world, len(all_keys), dtype=torch.long, device=dev)
            dist.all_gather_into_tensor(counts_t.view(-1), mine)
            counts = counts_t.t().tolist()  # [modality][rank]
        else:
            counts = [[int(c)] for c in mine.tolist()]
        sizes = [sum(c) for c in counts]
        total = sum(sizes)
        scale = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(1).contiguous()
        ctx.scale_shape, ctx.scale_dtype = logit_scale.shape, logit_scale.dtype
        feats = torch.zeros(total, dp, dtype=torch.bfloat16, device=dev)
        inv = torch.ones(total, dtype=torch.float32, device=dev)
        blocks = []  # (key, first global row, rows) of this rank's rows of every modality
        base = 0
        for k, cnt in zip(all_keys, counts):
            off = base + sum(cnt[:rank])
            n = cnt[rank]
            if n:
                src = by_key[k].detach().contiguous()
                rows = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
                check(lib.mmcl_pack_rows(h, ptr(src), _DTYPES[src.dtype], n, d, None, n,
                                         int(l2_normalize), ptr(rows),
                                         ptr(inv[off:off + n]) if l2_normalize else None, st),
                      "mmcl_pack_rows")
                feats[off:off + n, :d] = rows
            if world > 1 and sum(cnt):
                blk = feats[base:base + sum(cnt)]
                check(lib.mmcl_all_gather_rows(h, blk.data_ptr() + sum(cnt[:rank]) * dp * 2,
                                               ptr(blk), counts_array(cnt), dp * 2, st),
                      "mmcl_all_gather_rows")
                if l2_normalize:
                    iv = inv[base:base + sum(cnt)]
                    check(lib.mmcl_all_gather_rows(h, iv.data_ptr() + sum(cnt[:rank]) * 4, ptr(iv),
                                                   counts_array(cnt), 4, st),
                          "mmcl_all_gather_rows")
            blocks.append((k, off, n))
            base += sum(cnt)
        band = torch.tensor(band_ends(sizes), dtype=torch.int32, device=dev)
        idx_all = torch.arange(total, device=dev)
        npos_all = (band.long() - idx_all).clamp_min(1).double()
        # five sums per local row -> this rank's share of the loss and of d(loss)/d(scale)
        acc = torch.zeros(2, dtype=torch.float64, device=dev)
        for k, off, n in blocks:
            if n == 0:
                continue
            stats = torch.empty(5, n, dtype=torch.float32, device=dev)
            wsb = lib.mmcl_sim_bce_workspace_bytes(h, n, total, dp)
            ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
            q = feats[off:off + n]
            check(lib.mmcl_sim_bce_fwd(h, ptr(q), n, ptr(feats), total, dp, ptr(scale),
                                       ptr(inv[off:off + n]), ptr(inv), off, ptr(band), ptr(stats),
                                       ptr(ws), wsb, st), "mmcl_sim_bce_fwd")
            sA, sB, sC, sE, sF = stats.double()
            npos = npos_all[off:off + n]
            nneg = total - npos
            acc[0] += ((sB - sC) / npos + (sA - sB) / nneg).sum()
            acc[1] += ((sF - sC) / npos + (sE - sF) / nneg).sum()
        if world > 1:
            dist.all_reduce(acc)
        loss = (acc[0] / total).float()
        dscale = (acc[1] / total / scale[0].double()).float()
        ctx.state = (h, dev, feats, inv, band, scale, blocks, by_key, keys, d, dp, total,
                     l2_normalize, grad_factor)
        ctx.save_for_backward(dscale)
        return loss

    @staticmethod
    def backward(ctx, gout):
        from .loss import _stream

        lib = _lib.load()
        (dscale,) = ctx.saved_tensors
        (h, dev, feats, inv, band, scale, blocks, by_key, keys, d, dp, total, l2_normalize,
         grad_factor) = ctx.state
        st = _stream(dev)
        g32 = gout.detach().to(device=dev, dtype=torch.float32).contiguous()
        grads = {}
        for k, off, n in blocks:
            src = by_key.get(k)
            if src is None or not src.requires_grad:
                continue
            if n == 0 or grad_factor == 0:
                grads[k] = torch.zeros_like(src)
                continue
            dx = torch.zeros(n, dp, dtype=torch.float32, device=dev)
            x = feats[off:off + n]
            check(lib.mmcl_sim_bce_bwd(h, ptr(x), n, ptr(feats), total, dp, ptr(scale),
                                       ptr(inv[off:off + n]), ptr(inv), off, ptr(band), ptr(dx), st),
                  "mmcl_sim_bce_bwd")
            dacc = dx[:, :d].contiguous()
            s_c = src.detach().contiguous()
            out = torch.empty_like(s_c)
            partner = x[:, :d].contiguous()  # (positive-term operand; its coefficient is 0 here)
            check(lib.mmcl_unpack_grad(h, ptr(dacc), ptr(partner), None, n, d, float(grad_factor),
                                       0.0, ptr(g32), ptr(scale), ptr(s_c), _DTYPES[s_c.dtype],
                                       ptr(inv[off:off + n].contiguous()), int(l2_normalize), None,
                                       ptr(out), _DTYPES[s_c.dtype], 1, st), "mmcl_unpack_grad")
            grads[k] = out
        gs = None
        if ctx.needs_input_grad[1]:
            gs = (dscale * gout).reshape(ctx.scale_shape).to(ctx.scale_dtype)
        return (None, gs, *[grads.get(k) if by_key[k].requires_grad else None for k in keys])


def modality_alignment_loss(embeddings: dict, logit_scale: torch.Tensor, l2_normalize: bool,
                            local_loss: bool, gather_with_grad: bool) -> torch.Tensor:
    """``_compute_modality_alignment_loss`` on the (gathered) embeddings, fused."""
    world, _ = dist_info()
    if logit_scale.numel() != 1:
        raise ValueError("modality_alignment needs a scalar logit_scale (as in the reference)")
    keys = list(embeddings.keys())
    for k in keys:
        e = embeddings[k]
        if not isinstance(e, torch.Tensor) or not e.is_cuda:
            raise ValueError(f"embedding '{k}' is not a CUDA tensor: mmlearn_b200 has no CPU fallback")
        if e.dtype not in _DTYPES:
            raise ValueError(f"embedding '{k}' has unsupported dtype {e.dtype}")
    if world > 1:
        # key union over the ranks, sorted (contrastive.py:457-467): a rank that lacks a modality
        # still walks through every collective with zero rows of it
        gathered: list = [None] * world
        dist.all_gather_object(gathered, keys)
        all_keys = sorted(set().union(*[set(g_) for g_ in gathered]))
        grad_factor = world if gather_with_grad else (0 if local_loss else 1)
    else:
        all_keys = keys  # world_size == 1: all_embeddings is the caller's dict, in its own order
        grad_factor = 1
    meta = (keys, all_keys, bool(l2_normalize), grad_factor)
    return _FusedAlignmentFn.apply(meta, logit_scale, *[embeddings[k] for k in keys])
