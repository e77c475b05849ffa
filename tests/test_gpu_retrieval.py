"""GPU parity of mmlearn_b200.RetrievalRecallAtK (tcgen05 similarity mainloop with a counting
epilogue, mmcl_sim_rank_fwd) against the golden outputs of the unmodified reference metric and the
pinned oracle.  Hit indicators are integer work: bit-exact on the fixtures (bf16-representable
inputs, no ties); at larger sizes at most a 1e-3 fraction of queries may sit within fp32 rounding
of a tie and flip."""
import numpy as np
import pytest
import torch

from oracle.retrieval_oracle import recall_at_k_oracle
from tests.test_retrieval_oracle import load_recall, recall_cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("name", recall_cases())
def test_golden(name):
    from mmlearn_b200 import RetrievalRecallAtK

    z, xs, ys, idxs = load_recall(name)
    for k in z["ks"].tolist():
        for agg in ("mean", "min"):
            met = RetrievalRecallAtK(top_k=k, reduction="none", aggregation=agg)
            for x, y, i in zip(xs, ys, idxs):
                met.update(torch.from_numpy(x).to(DEV), torch.from_numpy(y).to(DEV),
                           torch.from_numpy(i).to(DEV))
            assert float(met.compute()) == pytest.approx(float(z[f"value_k{k}_{agg}"]), abs=1e-7)
        assert np.array_equal(met.hits().cpu().numpy(), z[f"hits_k{k}"])
    # one pass serves every k
    for k in z["ks"].tolist():
        assert np.array_equal(met.hits(k).cpu().numpy(), z[f"hits_k{k}"])
    met.reset()
    with pytest.raises(ValueError):
        met.ranks()


@pytest.mark.parametrize("n,d,mix", [(1000, 72, 0.3), (4096, 512, 0.2), (8192, 768, 0.15)])
def test_vs_oracle(n, d, mix):
    from mmlearn_b200 import RetrievalRecallAtK

    g = torch.Generator().manual_seed(5)
    x = torch.randn(n, d, generator=g).to(torch.bfloat16).float()
    y = (mix * x + torch.randn(n, d, generator=g)).to(torch.bfloat16).float()
    bs = 1024
    xs = [x[i:i + bs] for i in range(0, n, bs)]
    # indexes are relative to the batch start in update(): permute the positives inside batches
    ys, idxs = [], []
    for i in range(0, n, bs):
        xb = x[i:i + bs]
        p = torch.randperm(len(xb), generator=g)
        yb = torch.empty_like(xb)
        yb[p] = y[i:i + bs]
        ys.append(yb)
        idxs.append(p)
    met = RetrievalRecallAtK(top_k=10, reduction="none")
    for xb, yb, ib in zip(xs, ys, idxs):
        met.update(xb.to(DEV), yb.to(DEV), ib.to(DEV))
    ranks = met.ranks().cpu().numpy()
    for k in (1, 10, 100):
        value, hits, ref_ranks = recall_at_k_oracle(xs, ys, idxs, k)
        assert float(met.compute(k)) == pytest.approx(value, abs=2e-3)
        assert (met.hits(k).cpu().numpy() != hits).mean() <= 1e-3
    assert (np.abs(ranks - ref_ranks) > 0).mean() <= 2e-3
    assert np.abs(ranks - ref_ranks).max() <= 2


def test_errors_mirror_reference():
    from mmlearn_b200 import RetrievalRecallAtK

    with pytest.raises(ValueError, match="`top_k` has to be a positive integer or None"):
        RetrievalRecallAtK(top_k=0)
    with pytest.raises(ValueError, match="Expected argument `reduction`"):
        RetrievalRecallAtK(top_k=1, reduction="max")
    with pytest.raises(ValueError, match="Argument `aggregation`"):
        RetrievalRecallAtK(top_k=1, aggregation="sum")
    met = RetrievalRecallAtK(top_k=1, reduction="none")
    x = torch.randn(4, 8, device=DEV)
    with pytest.raises(ValueError, match="cannot be None"):
        met.update(x, x, None)
    with pytest.raises(RuntimeError, match="same shape"):
        met.update(x, x[:3], torch.arange(4, device=DEV))
    with pytest.raises(ValueError, match="2D tensors"):
        met.update(x[0], x[0], torch.arange(4, device=DEV))
    with pytest.raises(NotImplementedError):
        met(x, x, torch.arange(4, device=DEV))
