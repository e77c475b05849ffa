"""2-rank NCCL parity of the drop-in module against the golden outputs of the unmodified
reference run under 2-rank gloo (tests/golden/loss_w2_*.npz): all four
(local_loss, gather_with_grad) modes, uneven local batches, and a rank without local pairs."""
import os
import tempfile

import numpy as np
import pytest
import torch

from tests.golden_util import case_names, load_case

pytestmark = pytest.mark.gpu


def _worker(rank, world, name, init_file, ret, no_symm=False):
    import torch.distributed as dist

    if no_symm:
        os.environ["MMCL_NO_SYMM"] = "1"  # NCCL all-gathers instead of the peer-memory arena

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=dev)
    from mmlearn_b200 import ContrastiveLoss, LossPairSpec

    case = load_case(name)
    fn = ContrastiveLoss(**{"modality_alignment": False, **case["flags"]})
    leaves = {f"{m}_embedding": torch.from_numpy(v).to(dev).requires_grad_(True)
              for m, v in case["emb"][rank].items()}
    ids = {m: torch.from_numpy(v).to(dev) for m, v in case["ids"][rank].items()}
    s = torch.tensor(case["scale"], device=dev, dtype=torch.float32, requires_grad=True)
    loss = fn(leaves, ids, s, [LossPairSpec((a, b), w) for a, b, w in case["pairs"]])
    if loss.requires_grad:
        loss.backward()
    torch.cuda.synchronize()
    out = {"loss": float(loss.detach()),
           "dscale": float(s.grad) if s.grad is not None else 0.0}
    for k, v in leaves.items():
        out[f"grad_{k}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).float().cpu().numpy()
    if fn._labels:
        out["labels"] = next(iter(fn._labels.values())).cpu().numpy()
    ret[rank] = out
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", [n for n in case_names() if n.startswith("w2_")])
def test_golden_two_ranks(name):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    case = load_case(name)
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(2, name, os.path.join(td, "init"), ret), nprocs=2, join=True)
    for r in range(2):
        ref, got = case["out"][r], ret[r]
        assert got["loss"] == pytest.approx(float(ref["loss"]), rel=1e-3, abs=1e-6)
        for k in [k for k in got if k.startswith("grad_")]:
            g, e = torch.from_numpy(got[k]), torch.from_numpy(ref[k])
            if float(e.norm()) == 0:
                assert float(g.norm()) == 0
                continue
            assert float(g.norm()) == pytest.approx(float(e.norm()), rel=1e-2)
            assert float(torch.nn.functional.cosine_similarity(g.flatten(), e.flatten(), dim=0)) > 0.999
        # tiny problems (N=32, d=32): d(scale) nearly cancels, see tests/test_gpu_loss.py
        assert got["dscale"] == pytest.approx(float(ref["dscale"]), rel=5e-2, abs=2e-3)
        if "labels" in ref and "labels" in got:
            assert np.array_equal(got["labels"], ref["labels"])


def _worker_sequence(rank, world, names, init_file, ret):
    """One module instance fed a sequence of golden cases: exercises the cached-plan
    speculation (hit, miss with a different table, uneven ranks) under NCCL."""
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=dev)
    from mmlearn_b200 import ContrastiveLoss, LossPairSpec

    flags = load_case(names[0])["flags"]
    fn = ContrastiveLoss(**{"modality_alignment": False, **flags})
    outs = []
    for name in names:
        case = load_case(name)
        assert {k: v for k, v in case["flags"].items() if k != "cache_labels"} == \
            {k: v for k, v in flags.items() if k != "cache_labels"}
        leaves = {f"{m}_embedding": torch.from_numpy(v).to(dev).requires_grad_(True)
                  for m, v in case["emb"][rank].items()}
        ids = {m: torch.from_numpy(v).to(dev) for m, v in case["ids"][rank].items()}
        s = torch.tensor(case["scale"], device=dev, dtype=torch.float32, requires_grad=True)
        loss = fn(leaves, ids, s, [LossPairSpec((a, b), w) for a, b, w in case["pairs"]])
        if loss.requires_grad:
            loss.backward()
        torch.cuda.synchronize()
        out = {"loss": float(loss.detach())}
        for k, v in leaves.items():
            out[f"grad_{k}"] = (v.grad if v.grad is not None else torch.zeros_like(v)).float().cpu().numpy()
        out["steady_steps"] = fn.steady_steps
        outs.append(out)
    ret[rank] = outs
    dist.barrier()
    dist.destroy_process_group()


def test_speculation_sequence_two_ranks():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    # both cases were generated with l2_normalize, local_loss, gather_with_grad
    names = ["w2_ll1_gg1", "w2_ll1_gg1", "w2_uneven", "w2_uneven", "w2_ll1_gg1"]
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker_sequence, args=(2, names, os.path.join(td, "init"), ret), nprocs=2,
                 join=True)
    for r in range(2):
        for step, (name, got) in enumerate(zip(names, ret[r])):
            ref = load_case(name)["out"][r]
            assert got["loss"] == pytest.approx(float(ref["loss"]), rel=1e-3, abs=1e-6), (name, step)
            for k in [k for k in got if k.startswith("grad_")]:
                g, e = torch.from_numpy(got[k]), torch.from_numpy(ref[k])
                if float(e.norm()) == 0:
                    assert float(g.norm()) == 0
                    continue
                assert float(g.norm()) == pytest.approx(float(e.norm()), rel=1e-2), (name, step, k, r)


def test_nccl_transport_two_ranks():
    """The same golden case through the NCCL all-gather path (MMCL_NO_SYMM=1); the default path
    of the other tests is the symmetric peer-memory arena."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    name = "w2_uneven"
    case = load_case(name)
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(2, name, os.path.join(td, "init"), ret, True), nprocs=2, join=True)
    for r in range(2):
        ref, got = case["out"][r], ret[r]
        assert got["loss"] == pytest.approx(float(ref["loss"]), rel=1e-3, abs=1e-6)
        for k in [k for k in got if k.startswith("grad_")]:
            g, e = torch.from_numpy(got[k]), torch.from_numpy(ref[k])
            assert float(g.norm()) == pytest.approx(float(e.norm()), rel=1e-2)


def _check_sequence(names, ret, want_steady):
    for r in range(2):
        for step, (name, got) in enumerate(zip(names, ret[r])):
            ref = load_case(name)["out"][r]
            assert got["loss"] == pytest.approx(float(ref["loss"]), rel=1e-3, abs=1e-6), (name, step)
            for k in [k for k in got if k.startswith("grad_")]:
                g, e = torch.from_numpy(got[k]), torch.from_numpy(ref[k])
                if float(e.norm()) == 0:
                    assert float(g.norm()) == 0, (name, step, k)
                    continue
                assert float(g.norm()) == pytest.approx(float(e.norm()), rel=1e-2), (name, step, k)
                assert float(torch.nn.functional.cosine_similarity(
                    g.flatten(), e.flatten(), dim=0)) > 0.999, (name, step, k)
        assert [o["steady_steps"] for o in ret[r]] == want_steady, r


@pytest.mark.parametrize("name", ["w2_ll0_gg0", "w2_ll0_gg1", "w2_ll1_gg0", "w2_ll1_gg1"])
def test_native_step_two_ranks(name):
    """Steps 2 and 3 of an unchanged batch structure run through the native step executor
    (mmcl_step_*: peer stores into the symmetric arena, fences, no NCCL on the data path) and
    must reproduce the golden outputs of the reference in all four flag modes."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    names = [name] * 3
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker_sequence, args=(2, names, os.path.join(td, "init"), ret), nprocs=2,
                 join=True)
    _check_sequence(names, ret, [0, 1, 2])


def test_native_step_structure_change_on_one_rank():
    """Rank 1's batch loses its pairs while rank 0's stays the same: rank 0 alone could not
    notice, but the table every rank verifies holds every rank's fresh row, so both ranks leave
    the native path together (and return to it once the structure is steady again)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    names = ["w2_ll1_gg0", "w2_ll1_gg0", "w2_rank1_unpaired", "w2_rank1_unpaired", "w2_ll1_gg0",
             "w2_ll1_gg0", "w2_ll1_gg0"]
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker_sequence, args=(2, names, os.path.join(td, "init"), ret), nprocs=2,
                 join=True)
    _check_sequence(names, ret, [0, 1, 1, 1, 1, 2, 3])


def _worker_recall(rank, world, init_file, ret):
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=dev)
    from mmlearn_b200 import RetrievalRecallAtK

    xs, ys, idxs = _recall_data()
    met = RetrievalRecallAtK(top_k=5, reduction="none")
    for t in range(len(xs)):
        met.update(xs[t][rank].to(dev), ys[t][rank].to(dev), idxs[t][rank].to(dev))
    ret[rank] = {"value": {k: float(met.compute(k)) for k in (1, 5, 20)},
                 "hits5": met.hits().cpu().numpy(), "ranks": met.ranks().cpu().numpy()}
    dist.barrier()
    dist.destroy_process_group()


def _recall_data():
    """Two update steps with uneven per-rank batches; positives permuted inside each rank's batch."""
    g = torch.Generator().manual_seed(21)
    sizes = [[40, 56], [64, 24]]  # sizes[t][r]
    xs, ys, idxs = [], [], []
    for row in sizes:
        xt, yt, it = [], [], []
        for n in row:
            x = torch.randn(n, 64, generator=g).to(torch.bfloat16).float()
            y = (0.3 * x + torch.randn(n, 64, generator=g)).to(torch.bfloat16).float()
            p = torch.randperm(n, generator=g)
            yp = torch.empty_like(y)
            yp[p] = y
            xt.append(x)
            yt.append(yp)
            it.append(p)
        xs.append(xt)
        ys.append(yt)
        idxs.append(it)
    return xs, ys, idxs


def test_retrieval_recall_two_ranks():
    """Distributed RetrievalRecallAtK: every rank ranks its own queries against the all-gathered
    database; the global row order and index offsets are the reference's (update(),
    retrieval_recall.py:137-160), checked against the single-process oracle on the data laid out
    in that order."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from oracle.retrieval_oracle import recall_at_k_oracle

    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker_recall, args=(2, os.path.join(td, "init"), ret), nprocs=2, join=True)
    xs, ys, idxs = _recall_data()
    # what the reference's update() accumulates: per step, the ranks' batches concatenated, the
    # local indexes shifted by the rows of lower ranks (the oracle adds the rows of earlier steps)
    gx = [torch.cat(x) for x in xs]
    gy = [torch.cat(y) for y in ys]
    gi = [torch.cat([i + sum(len(q) for q in ib[:r]) for r, i in enumerate(ib)]) for ib in idxs]
    for k in (1, 5, 20):
        value, hits, ranks = recall_at_k_oracle(gx, gy, gi, k)
        for r in range(2):
            assert ret[r]["value"][k] == pytest.approx(value, abs=1e-6)
    value, hits, ranks = recall_at_k_oracle(gx, gy, gi, 5)
    # slice the oracle's global order (step-major, then rank) back into per-rank query order
    pos = 0
    per_rank = [[], []]
    for t in range(len(xs)):
        for r in range(2):
            n = len(xs[t][r])
            per_rank[r].append((hits[pos:pos + n], ranks[pos:pos + n]))
            pos += n
    for r in range(2):
        assert np.array_equal(ret[r]["hits5"], np.concatenate([h for h, _ in per_rank[r]]))
        assert np.array_equal(ret[r]["ranks"], np.concatenate([q for _, q in per_rank[r]]))


def _worker_rows(rank, world, rows, d, init_file, ret):
    import torch.distributed as dist

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"file://{init_file}", rank=rank, world_size=world,
                            device_id=dev)
    from mmlearn_b200 import ContrastiveLoss, LossPairSpec
    from oracle.contrastive_oracle import synthetic_inputs

    emb, ids = synthetic_inputs(rows * world, d, world=world)
    fn = ContrastiveLoss(l2_normalize=True, local_loss=True, gather_with_grad=True)
    outs = []
    for _ in range(3):
        leaves = {k: v.to(dev).requires_grad_(True) for k, v in emb[rank].items()}
        s = torch.tensor(1 / 0.07, device=dev, requires_grad=True)
        loss = fn(leaves, {k: torch.from_numpy(v).to(dev) for k, v in ids[rank].items()}, s,
                  [LossPairSpec(("rgb", "text"))])
        loss.backward()
        torch.cuda.synchronize()
        outs.append({"loss": float(loss.detach()), "dscale": float(s.grad),
                     "steady_steps": fn.steady_steps,
                     **{f"grad_{k}": v.grad.float().cpu().numpy() for k, v in leaves.items()}})
    ret[rank] = outs
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("rows,d,want_steady", [(18, 64, [0, 0, 0]), (264, 128, [0, 1, 2])])
def test_row_counts_around_the_fused_step_limits_two_ranks(rows, d, want_steady):
    """rows % 4 != 0: the fused multi-rank step (16-byte peer stores) refuses the shape and every
    rank stays on the general path; rows = 264 (not a multiple of the 256-key tile, so key tiles
    straddle the two ranks' blocks): native steps.  Both against the oracle."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp

    from oracle.contrastive_oracle import LossPairSpec as OPair
    from oracle.contrastive_oracle import contrastive_oracle, synthetic_inputs

    emb, ids = synthetic_inputs(rows * 2, d, world=2)
    ref = contrastive_oracle(emb, ids, 1 / 0.07, [OPair(("rgb", "text"))], l2_normalize=True,
                             local_loss=True, gather_with_grad=True, fast_match=True)
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker_rows, args=(2, rows, d, os.path.join(td, "init"), ret), nprocs=2, join=True)
    for r in range(2):
        assert [o["steady_steps"] for o in ret[r]] == want_steady
        for o in ret[r]:
            assert o["loss"] == pytest.approx(ref.losses[r], rel=1e-3)
            for k in ("rgb_embedding", "text_embedding"):
                g, e = torch.from_numpy(o[f"grad_{k}"]), ref.grads[r][k].float()
                assert float(g.norm()) == pytest.approx(float(e.norm()), rel=1e-2)
                assert float(torch.nn.functional.cosine_similarity(g.flatten(), e.flatten(), dim=0)) > 0.999
