"""TEST INFRASTRUCTURE ONLY -- writes ``tests/golden/zeroshot_*.npz`` by running the UNMODIFIED
``ZeroShotClassification.evaluation_step`` (mmlearn/tasks/zero_shot_classification.py:179-218,
loaded by ``oracle/ref_loader.load_reference_zero_shot``) on seeded embeddings with a capturing
metric collection: what it passes to ``metric_collection.update`` (logits, targets) is stored,
plus the top-k hit indicators ``torch.topk`` gives on those logits (what torchmetrics'
``Accuracy(top_k=k)`` counts).  Needs /root/reference:
    python oracle/make_golden_zero_shot.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference_zero_shot  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
KS = (1, 5)


class Capture:
    def __init__(self):
        self.logits, self.targets = [], []

    def update(self, logits, targets):
        self.logits.append(logits.detach().clone())
        self.targets.append(targets.detach().clone())


def run_reference(mod, queries, class_emb, targets):
    task = mod.ZeroShotClassification([mod.ClassificationTaskSpec("rgb", list(KS))], tokenizer=None)
    cap = Capture()
    task.metrics = {("rgb", 0): cap}
    task._embeddings_store = {0: class_emb}
    task.all_dataset_info = {0: {"num_classes": class_emb.shape[0]}}
    pl_module = types.SimpleNamespace(
        trainer=types.SimpleNamespace(sanity_checking=False),
        encode=lambda batch, modality: batch["emb"].clone())
    # a second dataset's rows are mixed in: evaluation_step must pick dataset 0's rows only
    n = queries.shape[0]
    ds = torch.zeros(n + 3, dtype=torch.long)
    ds[-3:] = 1
    emb = torch.cat([queries, torch.ones(3, queries.shape[1])])
    tg = torch.cat([targets, torch.zeros(3, dtype=torch.long)])
    task.evaluation_step(pl_module, {"dataset_index": ds, "emb": emb, "rgb_target": tg}, 0)
    assert len(cap.logits) == 1
    return cap.logits[0], cap.targets[0]


def main():
    mod = load_reference_zero_shot()
    rng = np.random.RandomState(5)

    def rows(n, d):
        return torch.from_numpy(rng.randn(n, d).astype(np.float32)).to(torch.bfloat16).float()

    cases = {"multiclass": (200, 37, 64), "wide": (96, 300, 128), "binary": (64, 2, 32)}
    for name, (n, c, d) in cases.items():
        protos = rows(c, d)
        targets = torch.from_numpy(rng.randint(0, c, size=n).astype(np.int64))
        queries = (0.25 * protos[targets] + rows(n, d)).to(torch.bfloat16).float()
        # the reference stores normalised class embeddings (zero_shot_classification.py:163-169);
        # rounded to bf16-representable values so that both sides see identical inputs
        class_emb = (protos / protos.norm(p=2, dim=-1, keepdim=True)).to(torch.bfloat16).float()
        logits, tg = run_reference(mod, queries, class_emb, targets)
        flat = dict(queries=queries.numpy(), class_emb=class_emb.numpy(), targets=tg.numpy(),
                    logits=logits.numpy(), ks=np.array(KS, np.int64))
        if c > 2:
            for k in KS:
                top = torch.topk(logits, k, dim=1).indices
                flat[f"hits_k{k}"] = (top == tg[:, None]).any(1).numpy()
        np.savez(os.path.join(OUT, f"zeroshot_{name}.npz"), **flat)
        print(name, logits.shape, {k: float(flat[f'hits_k{k}'].mean()) for k in KS} if c > 2 else "")


if __name__ == "__main__":
    main()
