"""CPU check of the formulation the kernels implement together with the
coefficient table of mmlearn_b200/plan.py: a dense fp64 evaluation of

    dA_i = s w [ sum_j ( P_row[i,j]/den_row + P_col[i,j]/den_col(owner(j)) ) B^_j - pos_coef B^_i ]

(P_row / P_col = row- / column-normalised softmax of the shared logit matrix S, only log-sum-exp
vectors cross ranks, no [N,d] gradient reduce-scatter) must reproduce the gradients the
UNMODIFIED reference produced through autograd + all-gather/reduce-scatter
(tests/golden/loss_w*.npz), in all four (local_loss, gather_with_grad) modes, with uneven ranks."""
import numpy as np
import pytest
import torch

from mmlearn_b200.plan import PairPlan
from tests.golden_util import load_case

CASES = ["w1_identity", "w2_ll0_gg0", "w2_ll0_gg1", "w2_ll1_gg0", "w2_ll1_gg1", "w2_uneven"]


def _normalize_backward(g, x, inv):
    xh = x * inv[:, None]
    return inv[:, None] * (g - xh * (xh * g).sum(1, keepdim=True))


@pytest.mark.parametrize("name", CASES)
def test_fused_formulation_matches_reference_gradients(name):
    case = load_case(name)
    W, f = case["world"], case["flags"]
    (ma, mb, w), s = case["pairs"][0], case["scale"]
    A = [torch.from_numpy(case["emb"][r][ma]).double() for r in range(W)]
    B = [torch.from_numpy(case["emb"][r][mb]).double() for r in range(W)]
    counts = [a.shape[0] for a in A]  # identity ids in these cases: every row is a pair
    ia = [1 / a.norm(dim=1).clamp_min(1e-12) if f["l2_normalize"] else torch.ones(len(a)).double()
          for a in A]
    ib = [1 / b.norm(dim=1).clamp_min(1e-12) if f["l2_normalize"] else torch.ones(len(b)).double()
          for b in B]
    Ah = torch.cat([a * i[:, None] for a, i in zip(A, ia)])
    Bh = torch.cat([b * i[:, None] for b, i in zip(B, ib)])
    S = s * Ah @ Bh.T                      # the one logit matrix both directions share
    lse_row_all = torch.logsumexp(S, 1)    # direction a of the rank that owns the row
    lse_col_all = torch.logsumexp(S, 0)    # direction b of the rank that owns the column
    owner = torch.repeat_interleave(torch.arange(W), torch.tensor(counts))
    for r in range(W):
        plan = PairPlan(counts=counts, rank=r, local_loss=f["local_loss"],
                        gather_with_grad=f["gather_with_grad"], weight=w)
        lo, hi = plan.offset, plan.offset + plan.b_local
        ref = case["out"][r]
        # ---- loss (contrastive.py:135-144) from the row statistics only ----
        pos = torch.diagonal(S)
        if plan.global_loss or W == 1:
            rows = slice(0, plan.n_total)
        else:
            rows = slice(lo, hi)
        loss = w / (2 * plan.loss_rows) * (lse_row_all[rows] + lse_col_all[rows] - 2 * pos[rows]).sum()
        assert float(loss) == pytest.approx(float(ref["loss"]), rel=1e-5)
        # ---- gradients of the local rows ----
        cden = plan.col_denominators()
        for side, (X, Yh, xinv, lse_r, lse_c, Sm, key) in enumerate((
                (A[r], Bh, ia[r], lse_row_all, lse_col_all, S, f"grad_{ma}_embedding"),
                (B[r], Ah, ib[r], lse_col_all, lse_row_all, S.T, f"grad_{mb}_embedding"))):
            Sl = Sm[lo:hi]                                   # this rank's row block
            G = torch.exp(Sl - lse_r[lo:hi, None]) / plan.row_denominator()
            if cden is not None:
                den = torch.tensor(cden, dtype=torch.float64)[owner if W > 1 else owner * 0]
                G = G + torch.exp(Sl - lse_c[None, :]) / den[None, :]
            g_hat = s * w * (G @ Yh - plan.positive_coefficient() * Yh[lo:hi])
            got = _normalize_backward(g_hat, X, xinv) if f["l2_normalize"] else g_hat
            np.testing.assert_allclose(got.numpy(), ref[key], rtol=2e-4, atol=2e-6)
