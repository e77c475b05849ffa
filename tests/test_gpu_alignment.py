"""GPU parity of the fused modality_alignment term (mmcl_sim_bce_fwd / mmcl_sim_bce_bwd through
mmlearn_b200/alignment.py) against a dense fp32 torch evaluation of the reference's formula
(contrastive.py:344-413) on the same bf16-rounded inputs, at sizes beyond the golden fixtures
(those run through ContrastiveLoss in tests/test_gpu_loss.py and tests/test_gpu_multirank.py).
Tolerances: loss 1e-3, gradient norms 1e-2 (cosine > 0.999), d(scale) 1e-2 relative."""
import pytest
import torch
import torch.nn.functional as F

from mmlearn_b200.alignment import alignment_targets

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _dense_reference(embs, scale, l2_normalize):
    """The reference's formula, dense, fp32 (leaves: fp32 copies of the inputs)."""
    leaves = {k: v.detach().float().clone().requires_grad_(True) for k, v in embs.items()}
    s = torch.tensor(float(scale), device=DEV, requires_grad=True)
    feats = torch.cat([F.normalize(v, dim=-1) if l2_normalize else v for v in leaves.values()], 0)
    target = alignment_targets([v.shape[0] for v in leaves.values()], DEV)
    logits = s * (feats @ feats.T)
    bce = F.binary_cross_entropy_with_logits(logits, target, reduction="none")
    pos = target > 0
    num_pos = target.sum(1)
    loss = (((bce * pos).sum(1) / num_pos) + ((bce * ~pos).sum(1) / (logits.shape[0] - num_pos))).mean()
    loss.backward()
    return float(loss), {k: v.grad for k, v in leaves.items()}, float(s.grad)


@pytest.mark.parametrize("sizes,d,scale,l2,dtype", [
    ((300, 300), 64, 14.285, True, torch.bfloat16),
    ((513, 257, 130), 160, 10.0, True, torch.float32),
    ((700, 1200), 512, 5.0, True, torch.bfloat16),
    ((260, 260), 128, 0.5, False, torch.float32),
])
def test_fused_alignment_against_dense_torch(sizes, d, scale, l2, dtype):
    from mmlearn_b200.alignment import modality_alignment_loss

    g = torch.Generator().manual_seed(17)
    names = ["rgb", "text", "audio"][:len(sizes)]
    base = torch.randn(max(sizes), d, generator=g)
    embs = {}
    for i, (nm, n) in enumerate(zip(names, sizes)):
        x = (0.5 * base[:n] + torch.randn(n, d, generator=g)) * (0.3 if not l2 else 1.0)
        embs[f"{nm}_embedding"] = x.to(torch.bfloat16).to(dtype).to(DEV).requires_grad_(True)
    s = torch.tensor(scale, device=DEV, requires_grad=True)
    loss = modality_alignment_loss(embs, s, l2, False, False)
    loss.backward()
    torch.cuda.synchronize()
    ref_loss, ref_grads, ref_ds = _dense_reference(embs, scale, l2)
    assert float(loss) == pytest.approx(ref_loss, rel=1e-3)
    for k, v in embs.items():
        got, want = v.grad.float(), ref_grads[k]
        assert float(got.norm()) == pytest.approx(float(want.norm()), rel=1e-2)
        assert float(F.cosine_similarity(got.flatten(), want.flatten(), dim=0)) > 0.999
    assert float(s.grad) == pytest.approx(ref_ds, rel=1e-2, abs=1e-6)


def test_frozen_modality_and_scale_buffer():
    from mmlearn_b200.alignment import modality_alignment_loss

    g = torch.Generator().manual_seed(3)
    a = torch.randn(200, 64, generator=g).to(DEV).requires_grad_(True)
    b = torch.randn(150, 64, generator=g).to(DEV)  # no gradient wanted
    loss = modality_alignment_loss({"rgb_embedding": a, "text_embedding": b},
                                   torch.tensor(8.0, device=DEV), True, False, False)
    loss.backward()
    assert a.grad is not None and b.grad is None and float(a.grad.norm()) > 0
