"""CHECKER ONLY -- chunked fp32 torch evaluation of one rank's loss, d(scale) and sampled gradient
rows at sizes where neither the reference nor ``oracle/`` can materialise the logits.

Used by ``bench.py``'s ``parity`` block (outside every timed region, at every GPU count) and by
``tests/``; the product never imports it.  It restates the reference semantics for the headline
flag set ``l2_normalize + local_loss + gather_with_grad`` with equal rows per rank
(mmlearn/modules/losses/contrastive.py:89-90 normalise, :196-212 labels, :321-332 logits,
:135-144 cross-entropies and pair weights, :506 gradient of the gathered copies) in the form of
SURVEY.md Appendix A:

    S        = s * A^ B^T                      (N x N, never held whole: row chunks)
    loss_r   = sum_p w_p/(2b) sum_{i in rank r} (lse_row[i] + lse_col[i] - 2 S_ii)
    dA^[i]   = s w/(2b) ( sum_j (softmax_row[i,j] + exp(S_ij - lse_col[j])) B^_j - 2 B^_i )
    dB^[i]   = s w/(2b) ( sum_j (exp(S_ji - lse_col[i]) + exp(S_ji - lse_row[j])) A^_j - 2 A^_i )
    dx       = (g - x^ (x^ . g)) / max(|x|, 1e-12)                 (normalise backward)

Inputs are the bf16-rounded embeddings of ALL ranks (every rank of the bench regenerates the
global synthetic batch from its seeds), evaluated in fp32 with TF32 off.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


@torch.no_grad()
def rank_check(embs: dict, pairs: list, rank: int, world: int, chunk: int = 4096,
               sample: int = 256, seed: int = 5) -> dict:
    """``embs``: name -> [N, d] tensor (any float dtype, on the GPU), the gathered raw embeddings
    in rank order.  ``pairs``: list of (name_a, name_b, weight, scale).  Returns this rank's
    ``loss``, ``dscale`` (list, one per pair), the sampled LOCAL row indices ``rows`` and
    ``grads`` (name -> fp32 [sample, d] gradient rows w.r.t. the raw embeddings)."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        any_t = next(iter(embs.values()))
        N, d = any_t.shape
        dev = any_t.device
        b = N // world
        off = rank * b
        raw = {k: v.float() for k, v in embs.items()}
        hat = {k: F.normalize(v, p=2, dim=-1) for k, v in raw.items()}
        rows = torch.randperm(b, generator=torch.Generator().manual_seed(seed))[:min(sample, b)]
        rows = rows.to(dev)
        gidx = rows + off
        ghat = {k: torch.zeros(rows.numel(), d, device=dev) for k in embs}
        loss = 0.0
        dscale = []
        for (na, nb, w, s) in pairs:
            A, B = hat[na], hat[nb]
            lse_row = torch.empty(N, device=dev)
            e_row = torch.empty(N, device=dev)
            cmax = torch.full((N,), -float("inf"), device=dev)
            csum = torch.zeros(N, device=dev)
            for i in range(0, N, chunk):
                S = s * (A[i:i + chunk] @ B.T)
                lr = torch.logsumexp(S, dim=1)
                lse_row[i:i + chunk] = lr
                e_row[i:i + chunk] = (torch.exp(S - lr[:, None]) * S).sum(1)
                m = S.max(dim=0).values
                nm = torch.maximum(cmax, m)
                csum = csum * torch.exp(cmax - nm) + torch.exp(S - nm[None, :]).sum(0)
                cmax = nm
            lse_col = cmax + torch.log(csum)
            e_col = torch.zeros(N, device=dev)
            for i in range(0, N, chunk):
                S = s * (A[i:i + chunk] @ B.T)
                e_col += (torch.exp(S - lse_col[None, :]) * S).sum(0)
            diag = s * (A * B).sum(1)
            loc = slice(off, off + b)
            loss += w / (2 * b) * float((lse_row[loc] + lse_col[loc] - 2 * diag[loc]).double().sum())
            dscale.append(w / (2 * b) * float(
                (e_row[loc] + e_col[loc] - 2 * diag[loc]).double().sum()) / s)
            # sampled gradient rows (w.r.t. the normalised rows), both modalities of the pair
            Sa = s * (A[gidx] @ B.T)                                     # rows gidx of S
            Ga = torch.exp(Sa - lse_row[gidx, None]) + torch.exp(Sa - lse_col[None, :])
            ghat[na] += s * w / (2 * b) * (Ga @ B - 2 * B[gidx])
            Sb = s * (B[gidx] @ A.T)                                     # columns gidx of S
            Gb = torch.exp(Sb - lse_col[gidx, None]) + torch.exp(Sb - lse_row[None, :])
            ghat[nb] += s * w / (2 * b) * (Gb @ A - 2 * A[gidx])
        grads = {}
        used = {n for p in pairs for n in p[:2]}
        for k in used:
            x, xh, g = raw[k][gidx], hat[k][gidx], ghat[k]
            nrm = x.norm(dim=1, keepdim=True).clamp_min(1e-12)
            grads[k] = (g - xh * (xh * g).sum(1, keepdim=True)) / nrm
        return {"loss": loss, "dscale": dscale, "rows": rows, "grads": grads}
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev


def compare(ref: dict, loss: float, dscale: list, grads: dict) -> dict:
    """Relative errors in the north-star's terms: loss, d(scale) and gradient norms over the
    sampled rows (``grads``: name -> full local gradient tensor [b, d])."""
    out = {"loss_rel": abs(loss - ref["loss"]) / max(abs(ref["loss"]), 1e-30)}
    out["dscale_rel"] = max(
        (abs(g - r) / max(abs(r), 1e-30) for g, r in zip(dscale, ref["dscale"])), default=0.0)
    gr, cs = 0.0, 1.0
    for k, want in ref["grads"].items():
        got = grads[k][ref["rows"]].float()
        gr = max(gr, abs(float(got.norm()) - float(want.norm())) / max(float(want.norm()), 1e-30))
        cs = min(cs, float(F.cosine_similarity(got.flatten(), want.flatten(), dim=0)))
    out["grad_rel"] = gr
    out["grad_cos_min"] = cs
    return out
