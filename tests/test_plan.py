"""CPU tests of the host logic (mmlearn_b200/plan.py): labels bit-exact with the reference's
golden outputs, the coefficient table, and the 2-rank count exchange over gloo."""
import os
import tempfile

import numpy as np
import pytest
import torch

from mmlearn_b200.plan import PairPlan, all_gather_ints, dist_info
from tests.golden_util import case_names, load_case


def test_single_process_plan():
    p = PairPlan(counts=[48], rank=0, local_loss=True, gather_with_grad=True)
    assert p.n_total == 48 and p.b_local == 48 and p.offset == 0
    assert torch.equal(p.labels(), torch.arange(48))
    assert p.labels().dtype == torch.long
    assert p.loss_rows == 48 and p.row_denominator() == 96.0
    assert p.col_denominators() == [96.0]
    assert p.positive_coefficient() == pytest.approx(1 / 48)
    assert dist_info() == (1, 0)


@pytest.mark.parametrize("name", [n for n in case_names() if n.startswith("w2_")])
def test_labels_match_reference_golden(name):
    """Labels cached by the unmodified reference (cache_labels=True cases) == PairPlan.labels."""
    case = load_case(name)
    counts = []
    for r in range(2):
        if "rgb" not in case["ids"][r] or "text" not in case["ids"][r]:
            counts.append(0)  # a rank that lacks a modality has no pair (contrastive.py:268-277)
            continue
        a, b = case["ids"][r]["rgb"], case["ids"][r]["text"]
        counts.append(int((a[:, None, :] == b[None, :, :]).all(-1).sum()))
    for r in range(2):
        plan = PairPlan(counts=counts, rank=r, local_loss=case["flags"]["local_loss"],
                        gather_with_grad=case["flags"]["gather_with_grad"])
        ref = case["out"][r]
        if "labels" in ref:
            assert np.array_equal(plan.labels().numpy(), ref["labels"])
        assert plan.n_total == sum(counts)


def test_mode_table():
    # (local_loss, gather_with_grad) -> what the rank-local gradient must be
    counts = [12, 20]
    tt = PairPlan(counts, 1, True, True)
    assert tt.loss_rows == 20 and tt.row_denominator() == 40.0
    assert tt.col_denominators() == [24.0, 40.0]
    assert tt.positive_coefficient() == pytest.approx(1 / 40 + 1 / 40)
    assert torch.equal(tt.labels(), torch.arange(12, 32))
    tf = PairPlan(counts, 0, True, False)
    assert tf.col_denominators() is None and tf.positive_coefficient() == pytest.approx(1 / 24)
    assert torch.equal(tf.labels(), torch.arange(0, 12))
    ft = PairPlan(counts, 0, False, True)  # W x d(global loss)
    assert ft.global_loss and ft.loss_rows == 32
    assert ft.row_denominator() == 32.0 and ft.col_denominators() == [32.0, 32.0]
    assert torch.equal(ft.labels(), torch.arange(32))
    ff = PairPlan(counts, 0, False, False)
    assert ff.row_denominator() == 64.0 and ff.positive_coefficient() == pytest.approx(1 / 32)
    # a rank that skipped the pair contributes no rows and gets empty labels
    sk = PairPlan([16, 0], 1, True, False)
    assert sk.b_local == 0 and sk.labels().numel() == 0
    # cross-rank duplicate ids: the reference's label slice is taken from arange(n_logits)
    dup = PairPlan([4, 4], 1, True, True, n_logits=10)
    assert dup.labels().tolist() == [4, 5, 6, 7, 8, 9]


def _gloo_worker(rank, world, init_file, ret):
    import torch.distributed as dist

    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    assert dist_info() == (world, rank)
    local = torch.tensor([10 + rank, rank, 7], dtype=torch.long)
    out = all_gather_ints(local, world)
    counts = [int(out[r, 0]) for r in range(world)]
    plan = PairPlan(counts, rank, True, True)
    ret[rank] = (out.tolist(), plan.offset, plan.labels().tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_count_exchange_two_ranks_gloo():
    import torch.multiprocessing as mp

    mgr = mp.Manager()
    ret = mgr.dict()
    with tempfile.TemporaryDirectory() as td:
        mp.spawn(_gloo_worker, args=(2, os.path.join(td, "init"), ret), nprocs=2, join=True)
    for r in range(2):
        table, off, labels = ret[r]
        assert table == [[10, 0, 7], [11, 1, 7]]
        assert off == (0 if r == 0 else 10)
        assert labels == (list(range(10)) if r == 0 else list(range(10, 21)))


def test_steady_state_table_detection():
    """mmlearn_b200.steady.table_rows: the native step executor is only armed when EVERY rank
    reports [rows, identity=1, rows, rows] for EVERY pair with the same rows > 0."""
    from mmlearn_b200.steady import table_rows

    ok = [[[64, 1, 64, 64], [64, 1, 64, 64]], [[64, 1, 64, 64], [64, 1, 64, 64]]]
    assert table_rows(ok, 2, 2) == 64
    for r, p, col, val in [(1, 0, 0, 63), (0, 1, 1, 0), (1, 1, 2, 65), (0, 0, 3, 60)]:
        bad = [[list(x) for x in row] for row in ok]
        bad[r][p][col] = val
        assert table_rows(bad, 2, 2) is None
    assert table_rows([[[0, 1, 0, 0]]], 1, 1) is None      # nobody has the pair
    assert table_rows([[[12, 1, 12, 12]], [[20, 1, 20, 20]]], 2, 1) is None  # uneven ranks
    assert table_rows([], 1, 0) is None
