"""The native step executor (mmcl_step_*, mmlearn_b200/steady.py): from the second call with an
unchanged batch structure on, forward and backward are one C call each.  Every such step must give
the numbers of the general path (same kernels, same order) and of the pinned oracle, must be
verified on the device, and must hand over to the general path when the structure changes."""
import numpy as np
import pytest
import torch

from oracle.contrastive_oracle import LossPairSpec as OPair
from oracle.contrastive_oracle import contrastive_oracle, synthetic_inputs
from tests.test_gpu_loss import _compare

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _step(fn, emb, ids, scale, pairs, dtype, grad_mask=None):
    from mmlearn_b200 import LossPairSpec

    leaves = {k: v.to(DEV).to(dtype).clone().requires_grad_(grad_mask is None or k in grad_mask)
              for k, v in emb.items()}
    s = torch.tensor(scale, device=DEV, dtype=torch.float32, requires_grad=True)
    loss = fn(leaves, {k: torch.from_numpy(v).to(DEV) for k, v in ids.items()}, s,
              [LossPairSpec((a, b), w) for a, b, w in pairs])
    loss.backward()
    torch.cuda.synchronize()
    return loss, leaves, s


@pytest.mark.parametrize("n,d,dtype,l2", [(512, 128, torch.bfloat16, True),
                                          (1000, 768, torch.float32, True),
                                          (384, 64, torch.float16, False),
                                          (4096, 512, torch.bfloat16, True)])
def test_steady_steps_match_general_path_and_oracle(n, d, dtype, l2):
    from mmlearn_b200 import ContrastiveLoss

    emb, ids = synthetic_inputs(n, d)
    scale = 1 / 0.07 if l2 else 0.05
    pairs = [("rgb", "text", 1.3)]
    flags = dict(l2_normalize=l2, local_loss=True, gather_with_grad=True)
    ref = contrastive_oracle(emb, ids, scale, [OPair((a, b), w) for a, b, w in pairs],
                             fast_match=True, **flags)
    fn = ContrastiveLoss(**flags)
    first = _step(fn, emb[0], ids[0], scale, pairs, dtype)  # general path, fills the caches
    assert fn.steady_steps == 0
    for k in range(1, 4):
        loss, leaves, s = _step(fn, emb[0], ids[0], scale, pairs, dtype)
        assert fn.steady_steps == k
        _compare(loss, leaves, s, ref.losses[0], ref.grads[0], ref.dscale[0], ref.dscale_gross[0])
        # same arithmetic as the general path, except that the fused step takes the positive logit
        # out of the tensor-core tile (another fp32 summation order than the dot-product kernel)
        assert float(loss) == pytest.approx(float(first[0]), rel=2e-5)
        for key in leaves:  # (fp32 atomics may reorder; a bf16 output may flip its last bit)
            g0, g1 = first[1][key].grad.float(), leaves[key].grad.float()
            assert float((g1 - g0).norm()) <= 2e-3 * float(g0.norm())
        assert float(s.grad) == pytest.approx(float(first[2].grad), rel=1e-4, abs=1e-7)


def test_steady_three_modalities_per_pair_scale_and_frozen_tower():
    from mmlearn_b200 import ContrastiveLoss, Modalities

    if not Modalities.has_modality("dna"):
        Modalities.register_modality("dna")
    emb, ids = synthetic_inputs(768, 256, modalities=("rgb", "dna", "text"))
    pairs = [("rgb", "dna", 1.0), ("rgb", "text", 0.5), ("dna", "text", 2.0)]
    scales = [1 / 0.07, 10.0, 50.0]
    ref = contrastive_oracle(emb, ids, scales, [OPair((a, b), w) for a, b, w in pairs],
                             l2_normalize=True, fast_match=True)
    fn = ContrastiveLoss(l2_normalize=True)
    for k in range(3):
        loss, leaves, s = _step(fn, emb[0], ids[0], scales, pairs, torch.bfloat16)
        assert fn.steady_steps == k
        _compare(loss, leaves, s, ref.losses[0], ref.grads[0], ref.dscale[0], ref.dscale_gross[0])
    # a frozen tower (text carries no gradient): new plan, same numbers for the other two
    before = fn.steady_steps
    mask = {"rgb_embedding", "dna_embedding"}
    for k in range(3):
        loss, leaves, s = _step(fn, emb[0], ids[0], scales, pairs, torch.bfloat16, grad_mask=mask)
        assert leaves["text_embedding"].grad is None
        sub = {key: leaves[key] for key in mask}
        _compare(loss, sub, s, ref.losses[0], ref.grads[0], ref.dscale[0], ref.dscale_gross[0])
    assert fn.steady_steps > before  # the later frozen-tower steps ran natively again


def test_steady_eval_then_train_and_structure_change():
    from mmlearn_b200 import ContrastiveLoss, LossPairSpec

    n, d = 640, 128
    emb, ids = synthetic_inputs(n, d)
    flags = dict(l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=True)
    fn = ContrastiveLoss(**flags)
    pairs = [("rgb", "text", 1.0)]
    ref = contrastive_oracle(emb, ids, 1 / 0.07, [OPair(("rgb", "text"))], fast_match=True,
                             **{k: v for k, v in flags.items() if k != "cache_labels"})
    e = {k: v.to(DEV).to(torch.bfloat16) for k, v in emb[0].items()}
    i = {k: torch.from_numpy(v).to(DEV) for k, v in ids[0].items()}
    with torch.no_grad():  # eval steps have their own plan (no state saved for backward)
        for k in range(3):
            loss = fn(e, i, torch.tensor(1 / 0.07, device=DEV), [LossPairSpec(("rgb", "text"))])
            assert float(loss) == pytest.approx(ref.losses[0], rel=1e-3)
    assert fn.steady_steps == 2
    for k in range(3):
        loss, leaves, s = _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)
        _compare(loss, leaves, s, ref.losses[0], ref.grads[0], ref.dscale[0], ref.dscale_gross[0])
    assert fn.steady_steps == 4
    # same shapes, different structure: the device-side verification must catch it
    rng = np.random.RandomState(3)
    ids2 = {k: v.copy() for k, v in ids[0].items()}
    ids2["text"] = ids2["text"][rng.permutation(n)]
    ids2["text"][:50, 1] += 100_000
    ref2 = contrastive_oracle([emb[0]], [ids2], 1 / 0.07, [OPair(("rgb", "text"))],
                              **{k: v for k, v in flags.items() if k != "cache_labels"})
    for k in range(2):
        loss, leaves, s = _step(fn, emb[0], ids2, 1 / 0.07, pairs, torch.bfloat16)
        _compare(loss, leaves, s, ref2.losses[0], ref2.grads[0], ref2.dscale[0],
                 ref2.dscale_gross[0])
    assert fn.steady_steps == 4  # permuted / unmatched ids are not the steady state
    # duplicate ids inside one modality: identity by position, but not a valid steady state
    ids3 = {k: v.copy() for k, v in ids[0].items()}
    ids3["rgb"][5] = ids3["rgb"][4]
    ids3["text"][5] = ids3["text"][4]
    ref3 = contrastive_oracle([emb[0]], [ids3], 1 / 0.07, [OPair(("rgb", "text"))],
                              **{k: v for k, v in flags.items() if k != "cache_labels"})
    _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)
    _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)
    assert fn.steady_steps == 5
    loss, leaves, s = _step(fn, emb[0], ids3, 1 / 0.07, pairs, torch.bfloat16)
    assert fn.steady_steps == 5
    _compare(loss, leaves, s, ref3.losses[0], ref3.grads[0], ref3.dscale[0], ref3.dscale_gross[0])


def test_two_forwards_before_backward():
    """Two evaluations in flight (e.g. two losses summed before one backward)."""
    from mmlearn_b200 import ContrastiveLoss, LossPairSpec

    emb, ids = synthetic_inputs(512, 128)
    flags = dict(l2_normalize=True, local_loss=True, gather_with_grad=True)
    ref = contrastive_oracle(emb, ids, 20.0, [OPair(("rgb", "text"))], fast_match=True, **flags)
    fn = ContrastiveLoss(**flags)
    i = {k: torch.from_numpy(v).to(DEV) for k, v in ids[0].items()}
    for k in range(3):
        a = {key: v.to(DEV).to(torch.bfloat16).requires_grad_(True) for key, v in emb[0].items()}
        b = {key: v.to(DEV).to(torch.bfloat16).requires_grad_(True) for key, v in emb[0].items()}
        s = torch.tensor(20.0, device=DEV, requires_grad=True)
        la = fn(a, i, s, [LossPairSpec(("rgb", "text"))])
        lb = fn(b, i, s, [LossPairSpec(("rgb", "text"))])
        (la + 2 * lb).backward()
        torch.cuda.synchronize()
        for key in a:
            r = torch.as_tensor(ref.grads[0][key]).float()
            assert float(a[key].grad.float().cpu().norm()) == pytest.approx(float(r.norm()), rel=1e-2)
            assert float(b[key].grad.float().cpu().norm()) == pytest.approx(2 * float(r.norm()), rel=1e-2)
        assert float(s.grad) == pytest.approx(3 * ref.dscale[0], rel=1e-2,
                                              abs=6e-3 * ref.dscale_gross[0])


def test_deferred_verification():
    """Opt-in ``deferred_verification``: forward() does not wait for the device; a valid step gives
    the reference's numbers, a step whose ids lost the assumed structure raises at the end of
    backward (before any optimizer could run) and the following call takes the general path."""
    from mmlearn_b200 import ContrastiveLoss

    n, d = 512, 128
    emb, ids = synthetic_inputs(n, d)
    flags = dict(l2_normalize=True, local_loss=True, gather_with_grad=True)
    ref = contrastive_oracle(emb, ids, 1 / 0.07, [OPair(("rgb", "text"))], fast_match=True, **flags)
    fn = ContrastiveLoss(**flags)
    fn.deferred_verification = True
    pairs = [("rgb", "text", 1.0)]
    for k in range(3):
        loss, leaves, s = _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)
        _compare(loss, leaves, s, ref.losses[0], ref.grads[0], ref.dscale[0], ref.dscale_gross[0])
    assert fn.steady_steps == 2
    ids2 = {k: v.copy() for k, v in ids[0].items()}
    ids2["text"] = ids2["text"][np.random.RandomState(3).permutation(n)]
    with pytest.raises(RuntimeError, match="deferred_verification"):
        _step(fn, emb[0], ids2, 1 / 0.07, pairs, torch.bfloat16)
    ref2 = contrastive_oracle([emb[0]], [ids2], 1 / 0.07, [OPair(("rgb", "text"))], **flags)
    loss, leaves, s = _step(fn, emb[0], ids2, 1 / 0.07, pairs, torch.bfloat16)  # general path
    _compare(loss, leaves, s, ref2.losses[0], ref2.grads[0], ref2.dscale[0], ref2.dscale_gross[0])
    # forward-only calls are checked at the next call
    e = {k: v.to(DEV).to(torch.bfloat16) for k, v in emb[0].items()}
    good = {k: torch.from_numpy(v).to(DEV) for k, v in ids[0].items()}
    bad = {k: torch.from_numpy(v).to(DEV) for k, v in ids2.items()}
    from mmlearn_b200 import LossPairSpec

    sc = torch.tensor(1 / 0.07, device=DEV)
    with torch.no_grad():
        fn(e, good, sc, [LossPairSpec(("rgb", "text"))])
        fn(e, good, sc, [LossPairSpec(("rgb", "text"))])
        fn(e, bad, sc, [LossPairSpec(("rgb", "text"))])  # native step on a batch that does not fit
        with pytest.raises(RuntimeError, match="deferred_verification"):
            fn(e, good, sc, [LossPairSpec(("rgb", "text"))])


def test_more_pairs_than_the_fused_step_handles_stay_on_the_general_path():
    """Four modalities / six pairs exceed the fused step's limits (4 pairs): the plug-in keeps
    using the general path (same numbers, no native steps)."""
    from mmlearn_b200 import ContrastiveLoss, Modalities

    for m in ("dna",):
        if not Modalities.has_modality(m):
            Modalities.register_modality(m)
    mods = ("rgb", "dna", "text", "audio")
    emb, ids = synthetic_inputs(256, 128, modalities=mods)
    pairs = [(a, b, 1.0 + 0.1 * i) for i, (a, b) in enumerate(
        (x, y) for k, x in enumerate(mods) for y in mods[k + 1:])]
    assert len(pairs) == 6
    ref = contrastive_oracle(emb, ids, 1 / 0.07, [OPair((a, b), w) for a, b, w in pairs],
                             l2_normalize=True, fast_match=True)
    fn = ContrastiveLoss(l2_normalize=True)
    for _ in range(3):
        loss, leaves, s = _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)
        _compare(loss, leaves, s, ref.losses[0], ref.grads[0], ref.dscale[0], ref.dscale_gross[0])
    assert fn.steady_steps == 0


def test_steady_steps_are_bitwise_reproducible():
    """The fused step writes every gradient element once (stores, plus scratch tiles summed in a
    fixed order) and reduces the loss in a fixed order, so repeated steps on the same batch must
    agree bit for bit.  This doubles as a race detector for the shared-memory hand-offs inside the
    backward kernel (conversion warps -> send warps -> partner CTA), which compute-sanitizer's
    racecheck cannot follow because they are ordered by mbarriers."""
    from mmlearn_b200 import ContrastiveLoss

    n, d = 8192, 512  # 64 x-tiles x 32 key tiles per (pair, side): several tiles per CTA pair
    emb, ids = synthetic_inputs(n, d)
    fn = ContrastiveLoss(l2_normalize=True, local_loss=True, gather_with_grad=True)
    pairs = [("rgb", "text", 1.0)]
    _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)  # general path
    loss0, leaves0, s0 = _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)
    assert fn.steady_steps == 1
    for k in range(2, 12):
        loss, leaves, s = _step(fn, emb[0], ids[0], 1 / 0.07, pairs, torch.bfloat16)
        assert fn.steady_steps == k
        assert torch.equal(loss, loss0) and torch.equal(s.grad, s0.grad)
        for key in leaves:
            assert torch.equal(leaves[key].grad, leaves0[key].grad), (k, key)
