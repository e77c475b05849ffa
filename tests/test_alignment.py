"""The modality_alignment branch (mmlearn_b200/alignment.py) is plain torch, so its arithmetic is
checked on the CPU: target construction bit-exact against the literal restatement of the
reference's Python loop, value and gradients against the golden outputs of the unmodified
reference (tests/golden/loss_w*_align_*.npz: total = contrastive + alignment; the contrastive
part is taken from the pinned oracle), single process and 2-rank gloo."""
import os
import tempfile

import numpy as np
import pytest
import torch

from mmlearn_b200.alignment import alignment_targets, modality_alignment_loss
from oracle.contrastive_oracle import LossPairSpec, contrastive_oracle
from tests.golden_util import case_names, load_case


def _literal_targets(sizes):
    import itertools

    n = sum(sizes)
    t = np.eye(n, dtype=np.float32)
    for idx, nk in enumerate(sizes):
        off = 0 if idx == 0 else sizes[idx - 1]
        for i, j in itertools.combinations(range(nk), 2):
            t[i + off, j + off] = 1
    return t


@pytest.mark.parametrize("sizes", [[4, 4], [1, 1], [5, 3], [3, 5], [6, 4, 3], [2, 7, 4], [3, 3, 3, 3],
                                   [1, 6, 2]])
def test_targets_bit_exact(sizes):
    assert np.array_equal(alignment_targets(sizes).numpy(), _literal_targets(sizes))


def _contrastive_part(case):
    emb = [{f"{m}_embedding": torch.from_numpy(v) for m, v in e.items()} for e in case["emb"]]
    f = case["flags"]
    return contrastive_oracle(emb, case["ids"], case["scale"],
                              [LossPairSpec((a, b), w) for a, b, w in case["pairs"]],
                              l2_normalize=f["l2_normalize"], local_loss=f["local_loss"],
                              gather_with_grad=f["gather_with_grad"])


def _check_rank(case, base, r, loss, leaves, s):
    ref = case["out"][r]
    assert float(loss) + base.losses[r] == pytest.approx(float(ref["loss"]), rel=1e-5)
    for k, v in leaves.items():
        g = (v.grad if v.grad is not None else torch.zeros_like(v)) + base.grads[r][k]
        np.testing.assert_allclose(g.numpy(), ref[f"grad_{k}"], rtol=2e-4, atol=2e-6)
    assert float(s.grad) + float(base.dscale[r]) == pytest.approx(float(ref["dscale"]), rel=2e-4,
                                                                  abs=1e-6)


@pytest.mark.parametrize("name", [n for n in case_names() if n.startswith("w1_align")])
def test_alignment_single_process(name):
    case = load_case(name)
    base = _contrastive_part(case)
    f = case["flags"]
    leaves = {f"{m}_embedding": torch.from_numpy(v).clone().requires_grad_(True)
              for m, v in case["emb"][0].items()}
    s = torch.tensor(case["scale"], dtype=torch.float32, requires_grad=True)
    loss = modality_alignment_loss(leaves, s, f["l2_normalize"], f["local_loss"],
                                   f["gather_with_grad"])
    loss.backward()
    _check_rank(case, base, 0, loss.detach(), leaves, s)


def _worker(rank, world, name, init_file, ret):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    case = load_case(name)
    f = case["flags"]
    leaves = {f"{m}_embedding": torch.from_numpy(v).clone().requires_grad_(True)
              for m, v in case["emb"][rank].items()}
    s = torch.tensor(case["scale"], dtype=torch.float32, requires_grad=True)
    loss = modality_alignment_loss(leaves, s, f["l2_normalize"], f["local_loss"],
                                   f["gather_with_grad"])
    loss.backward()
    ret[rank] = (float(loss.detach()),
                 {k: (v.grad if v.grad is not None else torch.zeros_like(v)).numpy()
                  for k, v in leaves.items()}, float(s.grad))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("name", [n for n in case_names() if n.startswith("w2_align")])
def test_alignment_two_ranks_gloo(name):
    import torch.multiprocessing as mp

    case = load_case(name)
    base = _contrastive_part(case)
    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_worker, args=(2, name, os.path.join(td, "init"), ret), nprocs=2, join=True)
    for r in range(2):
        loss, grads, ds = ret[r]
        ref = case["out"][r]
        assert loss + base.losses[r] == pytest.approx(float(ref["loss"]), rel=1e-5)
        for k, g in grads.items():
            np.testing.assert_allclose(g + base.grads[r][k].numpy(), ref[f"grad_{k}"], rtol=2e-4,
                                       atol=2e-6)
        assert ds + float(base.dscale[r]) == pytest.approx(float(ref["dscale"]), rel=2e-4, abs=1e-6)
