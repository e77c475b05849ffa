"""ContrastiveLossHead (mmlearn_b200/task.py): the task-side glue around the loss -- clamp + exp of
the log logit scale (one kernel, in-place clamp like contrastive_pretraining.py:488-489) and the
normalisation of raw encoder outputs inside the fused loss -- against the pinned oracle fed what
the reference task feeds its loss (normalised outputs, exp of the clamped parameter)."""
import math

import pytest
import torch

from oracle.contrastive_oracle import LossPairSpec as OPair
from oracle.contrastive_oracle import contrastive_oracle, synthetic_inputs
from tests.test_gpu_loss import _compare

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("init,max_scale", [(1 / 0.07, 100.0), (250.0, 100.0), (0.5, 100.0)])
def test_head_matches_reference_task_arithmetic(init, max_scale):
    from mmlearn_b200.task import ContrastiveLossHead

    # uncorrelated pairs: with correlated ones the loss at the clamp (s = 100) is ~1e-6 and
    # relative tolerances are meaningless
    emb, ids = synthetic_inputs(768, 256, correlated=False)
    s_eff = min(max(init, 1.0), max_scale)  # clamp(log s, 0, log max)
    # the reference task normalises in encode() and the loss normalises again: idempotent
    ref = contrastive_oracle(emb, ids, s_eff, [OPair(("rgb", "text"))], l2_normalize=True,
                             local_loss=True, gather_with_grad=True, fast_match=True)
    head = ContrastiveLossHead(init_logit_scale=init, max_logit_scale=max_scale).to(DEV)
    idt = {k: torch.from_numpy(v).to(DEV) for k, v in ids[0].items()}
    for step in range(3):  # general path, then native step executor
        leaves = {k: v.to(DEV).to(torch.bfloat16).requires_grad_(True) for k, v in emb[0].items()}
        head.log_logit_scale.grad = None
        loss = head(leaves, idt)
        loss.backward()
        torch.cuda.synchronize()
        # in-place clamp of the parameter, exactly like the task's no_grad clamp_
        assert float(head.log_logit_scale) == pytest.approx(math.log(s_eff), rel=1e-6)
        # d loss / d log_s = s * d loss / d s
        g = head.log_logit_scale.grad / s_eff

        class _S:  # adapter for _compare (expects a tensor with .grad)
            grad = g

        _compare(loss, leaves, _S, ref.losses[0], ref.grads[0], ref.dscale[0], ref.dscale_gross[0])
    # pre-normalised outputs give the same loss (drop-in for an unchanged task)
    with torch.no_grad():
        pre = {k: torch.nn.functional.normalize(v.to(DEV), dim=-1) for k, v in emb[0].items()}
        l2 = head(pre, idt)
    assert float(l2) == pytest.approx(ref.losses[0], rel=1e-3)


def test_head_buffer_scale_and_default_pairs():
    from mmlearn_b200.task import ContrastiveLossHead

    emb, ids = synthetic_inputs(256, 64, modalities=("rgb", "text", "audio"))
    head = ContrastiveLossHead(learnable_logit_scale=False, init_logit_scale=10.0).to(DEV)
    assert "log_logit_scale" in dict(head.named_buffers()) and not list(head.parameters())
    leaves = {k: v.to(DEV).requires_grad_(True) for k, v in emb[0].items()}
    loss = head(leaves, {k: torch.from_numpy(v).to(DEV) for k, v in ids[0].items()})
    loss.backward()
    pairs = [OPair(("rgb", "text")), OPair(("rgb", "audio")), OPair(("text", "audio"))]
    ref = contrastive_oracle(emb, ids, 10.0, pairs, l2_normalize=True, local_loss=True,
                             gather_with_grad=True, fast_match=True)
    assert float(loss) == pytest.approx(ref.losses[0], rel=1e-3)


def test_eval_mode_does_not_clamp_the_parameter():
    """The reference clamps ``log_logit_scale`` only in ``training_step``
    (contrastive_pretraining.py:488-489); evaluation just exponentiates it."""
    import math

    from mmlearn_b200.task import ContrastiveLossHead

    head = ContrastiveLossHead().to("cuda:0")
    with torch.no_grad():
        head.log_logit_scale.fill_(math.log(250.0))  # above max_logit_scale = 100
    head.eval()
    assert float(head.logit_scale()) == pytest.approx(250.0, rel=1e-5)
    assert float(head.log_logit_scale) == pytest.approx(math.log(250.0), rel=1e-6)
    head.train()
    assert float(head.logit_scale()) == pytest.approx(100.0, rel=1e-5)
    assert float(head.log_logit_scale) == pytest.approx(math.log(100.0), rel=1e-6)
