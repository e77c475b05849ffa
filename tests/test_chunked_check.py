"""The chunked fp32 checker behind ``bench.py``'s ``parity`` block (tools/chunked_check.py) against
the pinned oracle: same loss, d(scale) and gradient rows on every rank, W = 1, 2 and 4, with
several pairs, pair weights and per-pair scales (CPU)."""

import numpy as np
import pytest
import torch

from oracle.contrastive_oracle import LossPairSpec as OPair
from oracle.contrastive_oracle import contrastive_oracle, synthetic_inputs
from tools.chunked_check import compare, rank_check


@pytest.mark.parametrize("world,mods,weights,scales", [
    (1, ("rgb", "text"), [1.0], [1 / 0.07]),
    (2, ("rgb", "text"), [1.0], [1 / 0.07]),
    (4, ("rgb", "depth", "text"), [1.0, 0.5, 2.0], [1 / 0.07, 10.0, 50.0]),
])
def test_chunked_check_matches_oracle(world, mods, weights, scales):
    n, d = 96 * world, 64
    emb, ids = synthetic_inputs(n, d, modalities=mods, world=world)
    names = [(a, b) for i, a in enumerate(mods) for b in mods[i + 1:]]
    opairs = [OPair(p, w) for p, w in zip(names, weights)]
    scale = scales if len(scales) > 1 else scales[0]
    ref = contrastive_oracle(emb, ids, scale, opairs, l2_normalize=True, local_loss=True,
                             gather_with_grad=True)
    full = {m: torch.cat([e[f"{m}_embedding"] for e in emb], 0) for m in mods}
    pairs = [(a, b, w, s) for (a, b), w, s in zip(names, weights, scales * len(names))]
    for r in range(world):
        got = rank_check(full, pairs, r, world, chunk=40, sample=17)
        assert got["loss"] == pytest.approx(ref.losses[r], rel=1e-5)
        want_ds = np.atleast_1d(np.asarray(ref.dscale[r], dtype=np.float64))
        if len(scales) == 1:
            assert sum(got["dscale"]) == pytest.approx(float(want_ds.sum()), rel=2e-3, abs=1e-6)
        else:
            np.testing.assert_allclose(got["dscale"], want_ds, rtol=2e-3, atol=1e-6)
        rows = got["rows"]
        for m in mods:
            want = ref.grads[r][f"{m}_embedding"][rows]
            torch.testing.assert_close(got["grads"][m], want, rtol=2e-3, atol=1e-7)
        # compare() on the oracle's own numbers reports ~zero error
        c = compare(got, ref.losses[r], got["dscale"],
                    {m: ref.grads[r][f"{m}_embedding"] for m in mods})
        assert c["loss_rel"] < 1e-5 and c["grad_rel"] < 2e-3 and c["grad_cos_min"] > 0.9999
