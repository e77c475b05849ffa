"""TEST INFRASTRUCTURE ONLY -- writes ``tests/golden/recall_*.npz`` by running the UNMODIFIED
reference metric ``RetrievalRecallAtK`` (mmlearn/modules/metrics/retrieval_recall.py, loaded by
``oracle/ref_loader.load_reference_retrieval``) on seeded inputs.  Needs /root/reference:
    python oracle/make_golden_recall.py
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference_retrieval  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KS = (1, 5, 10)


def cases():
    rng = np.random.RandomState(11)

    def rows(n, d, base=None, mix=0.7):
        x = rng.randn(n, d).astype(np.float32)
        if base is not None:
            x = mix * base + x
        return torch.from_numpy(x).to(torch.bfloat16).float().numpy()  # bf16-representable

    out = {}
    # (a) the task's usage: several batches, indexes = arange(batch) (zero_shot_retrieval.py:110)
    xs = [rows(16, 8) for _ in range(3)]
    out["task_batches"] = dict(x=xs, y=[rows(16, 8, x) for x in xs], idx=[np.arange(16)] * 3)
    # (b) one batch, positives at permuted positions
    x = rows(64, 32)
    perm = rng.permutation(64)
    y = np.empty_like(x)
    y[perm] = rows(64, 32, x, 0.5)
    out["permuted"] = dict(x=[x], y=[y], idx=[perm])
    # (c) ragged last batch, weakly correlated pairs (most ranks > 10)
    xs = [rows(n, 64) for n in (128, 128, 44)]
    out["ragged"] = dict(x=xs, y=[rows(len(x), 64, x, 0.25) for x in xs],
                         idx=[np.arange(len(x)) for x in xs])
    return out


def main():
    mod = load_reference_retrieval()
    for name, c in cases().items():
        flat = {"nb": np.int64(len(c["x"])), "ks": np.array(KS, np.int64)}
        for b, (x, y, i) in enumerate(zip(c["x"], c["y"], c["idx"])):
            flat[f"x{b}"], flat[f"y{b}"], flat[f"idx{b}"] = x, y, np.asarray(i, np.int64)
        for k in KS:
            for agg in ("mean", "min"):
                met = mod.RetrievalRecallAtK(top_k=k, reduction="none", aggregation=agg)
                for x, y, i in zip(c["x"], c["y"], c["idx"]):
                    met.update(torch.from_numpy(x), torch.from_numpy(y), torch.from_numpy(np.asarray(i)))
                flat[f"value_k{k}_{agg}"] = np.float64(float(met.compute()))
            # per-query hits through the reference's own _recall_at_k on the full problem
            xx = torch.cat([torch.from_numpy(x) for x in c["x"]])
            yy = torch.cat([torch.from_numpy(y) for y in c["y"]])
            off = np.cumsum([0] + [len(x) for x in c["x"]])[:-1]
            ii = torch.cat([torch.from_numpy(np.asarray(i)) + int(o) for i, o in zip(c["idx"], off)])
            xx = xx / xx.norm(dim=-1, keepdim=True)
            yy = yy / yy.norm(dim=-1, keepdim=True)
            sim = xx @ yy.T
            pos = torch.zeros_like(sim, dtype=torch.bool)
            pos[torch.arange(len(sim)), ii] = True
            flat[f"hits_k{k}"] = (mod._recall_at_k(sim, pos, k) > 0).numpy()
        np.savez_compressed(os.path.join(OUT, f"recall_{name}.npz"), **flat)
        print(name, {k: float(flat[f"value_k{k}_mean"]) for k in KS})


if __name__ == "__main__":
    main()
