"""GPU parity of the fused zero-shot scoring step (mmlearn_b200/zero_shot.py) against the golden
outputs of the unmodified reference ``evaluation_step`` and the oracle at a larger size.
Tolerance: scores |err| <= 2e-3 * (1 + |score|) (bf16 rows, fp32 accumulation, fp32 normalisation
factors; scores are 100 * cos); top-k hit indicators equal except where the reference's own scores
of the target and a competitor are closer than that tolerance."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.zero_shot_oracle import topk_hits, zero_shot_logits

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "zeroshot_*.npz")))


def _check(q, c, targets, ref_logits, ks):
    from mmlearn_b200 import ZeroShotScores

    zs = ZeroShotScores(c.to(DEV))
    got = zs.logits(q.to(DEV)).cpu()
    assert got.shape == ref_logits.shape
    tol = 2e-3 * (1 + ref_logits.abs())
    assert bool(((got - ref_logits).abs() <= tol).all())
    if c.shape[0] > 2:
        ranks = zs.target_ranks(q.to(DEV), targets.to(DEV)).cpu()
        # reference rank of the target: classes scoring strictly above it
        tgt = ref_logits.gather(1, targets[:, None])
        ref_rank = (ref_logits > tgt).sum(1)
        margin = (ref_logits - tgt).abs()
        margin.scatter_(1, targets[:, None], float("inf"))
        clear = margin.min(1).values > 2 * 2e-3 * (1 + tgt.abs().squeeze(1))
        assert bool((ranks[clear] == ref_rank[clear]).all())
        assert float(clear.float().mean()) > 0.7  # (300 classes: many near-ties with the target)
        for k in ks:
            hits = zs.topk_correct(None, None, int(k), ranks=ranks)
            want = topk_hits(ref_logits, targets, int(k))
            assert bool((hits[clear] == want[clear]).all())


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_golden(path):
    c = np.load(path)
    _check(torch.from_numpy(c["queries"]), torch.from_numpy(c["class_emb"]),
           torch.from_numpy(c["targets"]), torch.from_numpy(c["logits"]), c["ks"])


def test_imagenet_sized_against_oracle():
    g = torch.Generator().manual_seed(3)
    n, ncls, d = 4096, 1000, 512
    protos = torch.randn(ncls, d, generator=g)
    targets = torch.randint(0, ncls, (n,), generator=g)
    q = (0.2 * protos[targets] + torch.randn(n, d, generator=g)).to(torch.bfloat16).float()
    c = (protos / protos.norm(dim=1, keepdim=True)).to(torch.bfloat16).float()
    _check(q, c, targets, zero_shot_logits(q, c), (1, 5))


def test_errors():
    from mmlearn_b200 import ZeroShotScores

    with pytest.raises(ValueError):
        ZeroShotScores(torch.zeros(4, 16))  # CPU tensor: no fallback
    zs = ZeroShotScores(torch.randn(4, 16, device=DEV))
    with pytest.raises(ValueError):
        zs.logits(torch.randn(3, 8, device=DEV))
    assert zs.logits(torch.randn(0, 16, device=DEV)).shape == (0, 4)
