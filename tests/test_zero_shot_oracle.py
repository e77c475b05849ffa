"""The zero-shot scoring oracle against the golden outputs of the unmodified reference
``evaluation_step`` (tests/golden/zeroshot_*.npz, oracle/make_golden_zero_shot.py); CPU."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle.zero_shot_oracle import topk_hits, zero_shot_logits

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(glob.glob(os.path.join(GOLDEN, "zeroshot_*.npz")))


@pytest.mark.parametrize("path", CASES, ids=[os.path.basename(p)[:-4] for p in CASES])
def test_oracle_matches_reference(path):
    c = np.load(path)
    logits = zero_shot_logits(torch.from_numpy(c["queries"]), torch.from_numpy(c["class_emb"]))
    np.testing.assert_allclose(logits.numpy(), c["logits"], rtol=1e-5, atol=1e-5)
    for k in c["ks"]:
        key = f"hits_k{int(k)}"
        if key in c.files:
            assert np.array_equal(
                topk_hits(logits, torch.from_numpy(c["targets"]), int(k)).numpy(), c[key])


def test_fixtures_exist():
    assert len(CASES) == 3
