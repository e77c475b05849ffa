"""Pins oracle/retrieval_oracle.py against the golden outputs of the unmodified reference metric
(tests/golden/recall_*.npz, oracle/make_golden_recall.py) and checks the host-side row bookkeeping
of the distributed metric (mmlearn_b200.retrieval.global_layout) against a literal replay of the
reference's update() arithmetic (retrieval_recall.py:137-160)."""
import glob
import os

import numpy as np
import pytest
import torch

from mmlearn_b200.retrieval import global_layout
from oracle.retrieval_oracle import recall_at_k_oracle
from tests.golden_util import GOLDEN


def recall_cases():
    return sorted(os.path.basename(p)[len("recall_"):-4]
                  for p in glob.glob(os.path.join(GOLDEN, "recall_*.npz")))


def load_recall(name):
    z = np.load(os.path.join(GOLDEN, f"recall_{name}.npz"))
    nb = int(z["nb"])
    return z, [z[f"x{b}"] for b in range(nb)], [z[f"y{b}"] for b in range(nb)], \
        [z[f"idx{b}"] for b in range(nb)]


@pytest.mark.parametrize("name", recall_cases())
def test_oracle_matches_reference_golden(name):
    z, xs, ys, idxs = load_recall(name)
    for k in z["ks"].tolist():
        for agg in ("mean", "min"):
            value, hits, ranks = recall_at_k_oracle(xs, ys, idxs, k, agg)
            assert value == pytest.approx(float(z[f"value_k{k}_{agg}"]), abs=1e-7)
        assert np.array_equal(hits, z[f"hits_k{k}"])
        assert np.array_equal(ranks < k, hits)  # no exact ties in the fixtures


@pytest.mark.parametrize("sizes", [[[4, 4]], [[3, 5], [2, 2]], [[1, 2, 3], [4, 5, 6], [7, 8, 9]],
                                   [[8], [8], [3]]])
def test_global_layout_replays_reference_update(sizes):
    world = len(sizes[0])
    # literal replay of retrieval_recall.py:137-160 for every rank
    num_samples = 0
    want = [[] for _ in range(world)]
    for row in sizes:
        idx = [torch.arange(n) + num_samples for n in row]  # :137 on each rank
        cat = torch.cat(idx)  # :142-144 gather
        cum = np.cumsum(row)  # :150
        for dev in range(1, world):
            cat[cum[dev - 1]:cum[dev]] += cum[dev - 1]  # :151-154
        num_samples += int(cum[-1])  # :157
        lo = 0
        for r, n in enumerate(row):
            want[r].append(cat[lo:lo + n])
            lo += n
    for r in range(world):
        offsets, order = global_layout(sizes, r)
        for t, row in enumerate(sizes):
            assert torch.equal(torch.arange(row[r]) + offsets[t], want[r][t])
        assert order == [(rr, t) for t in range(len(sizes)) for rr in range(world)]
