"""TEST INFRASTRUCTURE ONLY -- generates ``tests/golden/*.npz`` by running the UNMODIFIED
reference loss (``/root/reference/mmlearn/modules/losses/contrastive.py``, loaded through
``oracle/ref_loader.py``) on small seeded inputs: single process and 2-rank gloo.

Run from the repo root, in the authoring container (needs /root/reference):
    python oracle/make_golden.py
The fixtures are committed; the GPU box never runs this script.
"""

from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.ref_loader import load_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ids(ds, ex):
    return np.stack([np.asarray(ds, np.int64), np.asarray(ex, np.int64)], 1)


def build_cases():
    """name -> dict(world, mods, emb[rank][mod], ids[rank][mod], scale, pairs, flags)."""
    cases = {}
    rng = np.random.RandomState(7)

    def emb(n, d, base=None):
        x = rng.randn(n, d).astype(np.float32)
        if base is not None:
            x = 0.5 * base + x
        return torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()

    # 1. single process, identity ids
    n, d = 48, 64
    a = emb(n, d)
    b = emb(n, d, a)
    cases["w1_identity"] = dict(
        world=1, emb=[{"rgb": a, "text": b}], ids=[{"rgb": _ids(np.zeros(n), np.arange(n)),
                                                    "text": _ids(np.zeros(n), np.arange(n))}],
        scale=1 / 0.07, pairs=[("rgb", "text", 1.0)],
        flags=dict(l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=True))
    # 2. single process, permuted + partial overlap + duplicate ids, no l2-normalise
    ida = _ids([0] * 40, list(range(40)))
    exb = list(range(10, 46))
    perm = rng.permutation(len(exb))
    exb = [exb[i] for i in perm] + [12, 12]  # duplicates in the second modality
    idb = _ids([0] * len(exb), exb)
    a = emb(40, 32)
    b = emb(len(exb), 32)
    cases["w1_ragged_ids"] = dict(
        world=1, emb=[{"rgb": a * 0.2, "text": b * 0.2}], ids=[{"rgb": ida, "text": idb}],
        scale=5.0, pairs=[("rgb", "text", 0.7)],
        flags=dict(l2_normalize=False, local_loss=False, gather_with_grad=False, cache_labels=False))
    # 3. three modalities, weights, one scale
    n, d = 24, 32
    a = emb(n, d)
    b = emb(n, d, a)
    c = emb(n, d, a)
    idn = _ids(np.zeros(n), np.arange(n))
    cases["w1_three_mods"] = dict(
        world=1, emb=[{"rgb": a, "text": b, "audio": c}],
        ids=[{"rgb": idn, "text": idn, "audio": idn}], scale=10.0,
        pairs=[("rgb", "text", 1.0), ("rgb", "audio", 0.5), ("text", "audio", 2.0)],
        flags=dict(l2_normalize=True, local_loss=False, gather_with_grad=False, cache_labels=False))
    # 4. two ranks, all four flag combinations
    n, d, W = 32, 32, 2
    a = emb(n, d)
    b = emb(n, d, a)
    for ll in (True, False):
        for gg in (True, False):
            embs, idss = [], []
            for r in range(W):
                sl = slice(r * n // W, (r + 1) * n // W)
                idr = _ids(np.zeros(n // W), np.arange(n)[sl])
                embs.append({"rgb": a[sl], "text": b[sl]})
                idss.append({"rgb": idr, "text": idr})
            cases[f"w2_ll{int(ll)}_gg{int(gg)}"] = dict(
                world=W, emb=embs, ids=idss, scale=1 / 0.07, pairs=[("rgb", "text", 1.0)],
                flags=dict(l2_normalize=True, local_loss=ll, gather_with_grad=gg, cache_labels=True))
    # 5. two ranks, uneven local batch (12 + 20), local loss with grad
    cut = 12
    embs = [{"rgb": a[:cut], "text": b[:cut]}, {"rgb": a[cut:], "text": b[cut:]}]
    idss = [{"rgb": _ids(np.zeros(cut), np.arange(cut)), "text": _ids(np.zeros(cut), np.arange(cut))},
            {"rgb": _ids(np.zeros(n - cut), np.arange(cut, n)),
             "text": _ids(np.zeros(n - cut), np.arange(cut, n))}]
    cases["w2_uneven"] = dict(
        world=2, emb=embs, ids=idss, scale=20.0, pairs=[("rgb", "text", 1.5)],
        flags=dict(l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=False))
    # 6. two ranks, rank 1 has no matching pair locally (skip_flag path, contrastive.py:283-301).
    # gather_with_grad must be False here: with True the reference itself deadlocks, because
    # rank 1 returns a constant 0.0 (no backward) while rank 0's backward waits in the
    # reduce-scatter of the autograd all-gather.
    embs = [{"rgb": a[:16], "text": b[:16]}, {"rgb": a[16:], "text": b[16:]}]
    idss = [{"rgb": _ids(np.zeros(16), np.arange(16)), "text": _ids(np.zeros(16), np.arange(16))},
            {"rgb": _ids(np.zeros(16), np.arange(16, 32)),
             "text": _ids(np.ones(16), np.arange(16, 32))}]
    cases["w2_rank1_unpaired"] = dict(
        world=2, emb=embs, ids=idss, scale=1 / 0.07, pairs=[("rgb", "text", 1.0)],
        flags=dict(l2_normalize=True, local_loss=True, gather_with_grad=False, cache_labels=False))
    # 6b. two ranks, rank 1 does not have the text modality at all (placeholder path of
    # _gather_dicts, contrastive.py:471-480,495); gather_with_grad=False for the reason above
    embs = [{"rgb": a[:16], "text": b[:16]}, {"rgb": a[16:]}]
    idss = [{"rgb": _ids(np.zeros(16), np.arange(16)), "text": _ids(np.zeros(16), np.arange(16))},
            {"rgb": _ids(np.zeros(16), np.arange(16, 32))}]
    cases["w2_rank1_missing_modality"] = dict(
        world=2, emb=embs, ids=idss, scale=1 / 0.07, pairs=[("rgb", "text", 1.0)],
        flags=dict(l2_normalize=True, local_loss=True, gather_with_grad=False, cache_labels=False))
    # 7. modality_alignment term (contrastive.py:344-413).  Unequal modality sizes in the
    # three-modality case exercise the reference's block-offset rule.
    n, d = 12, 32
    a = emb(n, d)
    b = emb(n, d, a)
    idn = _ids(np.zeros(n), np.arange(n))
    cases["w1_align_two_mods"] = dict(
        world=1, emb=[{"rgb": a, "text": b}], ids=[{"rgb": idn, "text": idn}], scale=10.0,
        pairs=[("rgb", "text", 1.0)],
        flags=dict(l2_normalize=True, local_loss=False, gather_with_grad=False, cache_labels=False,
                   modality_alignment=True))
    c = emb(5, d, a[:5])
    cases["w1_align_three_mods"] = dict(
        world=1, emb=[{"rgb": a[:10], "text": b[:7], "audio": c}],
        ids=[{"rgb": idn[:10], "text": idn[:7], "audio": idn[:5]}], scale=5.0,
        pairs=[("rgb", "text", 1.0), ("text", "audio", 0.5)],
        flags=dict(l2_normalize=True, local_loss=False, gather_with_grad=False, cache_labels=False,
                   modality_alignment=True))
    n, W = 16, 2
    a = emb(n, d)
    b = emb(n, d, a)
    for gg in (True, False):
        embs, idss = [], []
        for r in range(W):
            sl = slice(r * n // W, (r + 1) * n // W)
            idr = _ids(np.zeros(n // W), np.arange(n)[sl])
            embs.append({"rgb": a[sl], "text": b[sl]})
            idss.append({"rgb": idr, "text": idr})
        cases[f"w2_align_ll0_gg{int(gg)}"] = dict(
            world=W, emb=embs, ids=idss, scale=8.0, pairs=[("rgb", "text", 1.0)],
            flags=dict(l2_normalize=True, local_loss=False, gather_with_grad=gg,
                       cache_labels=False, modality_alignment=True))
    return cases


def _run_rank(rank, world, case, init_file, ret):
    ref = load_reference()
    if world > 1:
        dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank,
                                world_size=world)
    torch.manual_seed(0)
    Loss = ref["contrastive"].ContrastiveLoss
    Pair = ref["task_stub"].LossPairSpec
    loss_fn = Loss(**{"modality_alignment": False, **case["flags"]})
    leaves = {f"{m}_embedding": torch.tensor(v, requires_grad=True)
              for m, v in case["emb"][rank].items()}
    ids = {m: torch.from_numpy(v) for m, v in case["ids"][rank].items()}
    s = torch.tensor(float(case["scale"]), requires_grad=True)
    pairs = [Pair(modalities=(p[0], p[1]), weight=p[2]) for p in case["pairs"]]
    loss = loss_fn(leaves, ids, s, pairs)
    if loss.requires_grad:
        loss.backward()
    out = {"loss": loss.detach().double().numpy(),
           "dscale": (s.grad if s.grad is not None else torch.zeros(())).double().numpy()}
    for k, v in leaves.items():
        out[f"grad_{k}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
    if case["flags"].get("cache_labels") and loss_fn._labels:
        out["labels"] = next(iter(loss_fn._labels.values())).numpy()
    ret[rank] = out
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_case(case):
    world = case["world"]
    if world == 1:
        ret = {}
        _run_rank(0, 1, case, None, ret)
        return ret
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_run_rank, args=(world, case, os.path.join(td, "init"), ret), nprocs=world,
                 join=True)
    return dict(ret)


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = load_reference()
    only = [a.split("=", 1)[1] for a in sys.argv[1:] if a.startswith("--only=")]
    for name, case in build_cases().items():
        if only and not any(name.startswith(p) for p in only):
            continue  # keep the committed fixtures of the other cases untouched
        ret = run_case(case)
        flat = {"world": np.int64(case["world"]), "scale": np.float64(case["scale"]),
                "pairs": np.array([f"{p[0]}|{p[1]}|{p[2]!r}" for p in case["pairs"]]),
                "flags": np.array([f"{k}={int(v)}" for k, v in case["flags"].items()])}
        for r in range(case["world"]):
            for m, v in case["emb"][r].items():
                flat[f"r{r}_emb_{m}"] = v
            for m, v in case["ids"][r].items():
                flat[f"r{r}_ids_{m}"] = v
            for k, v in ret[r].items():
                flat[f"r{r}_out_{k}"] = v
        np.savez_compressed(os.path.join(OUT, f"loss_{name}.npz"), **flat)
        print(name, [float(ret[r]["loss"]) for r in range(case["world"])])
    if only:
        return
    # find_matching_indices: the reference function on seeded random ids (duplicates, misses)
    fmi = ref["example"].find_matching_indices
    rng = np.random.RandomState(3)
    flat = {}
    for i, (n, m, hi) in enumerate([(5, 5, 3), (17, 9, 4), (64, 80, 30), (1, 7, 2), (33, 33, 1000)]):
        a = np.stack([rng.randint(0, 2, n), rng.randint(0, hi, n)], 1).astype(np.int64)
        b = np.stack([rng.randint(0, 2, m), rng.randint(0, hi, m)], 1).astype(np.int64)
        ia, ib = fmi(torch.from_numpy(a), torch.from_numpy(b))
        flat[f"c{i}_a"], flat[f"c{i}_b"] = a, b
        flat[f"c{i}_ia"], flat[f"c{i}_ib"] = ia.numpy(), ib.numpy()
    np.savez_compressed(os.path.join(OUT, "match_random.npz"), **flat)
    print("wrote", sorted(os.listdir(OUT)))


if __name__ == "__main__":
    main()
