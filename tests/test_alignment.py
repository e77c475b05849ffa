"""Host logic of the modality_alignment term (mmlearn_b200/alignment.py), checked on the CPU: the
per-row positive band the kernels evaluate from tile coordinates reproduces, bit for bit, the
dense target matrix of the reference's Python loop (contrastive.py:375-391, block-offset quirk
included).  The arithmetic itself runs in the sm_100a kernels and is checked against the golden
outputs of the unmodified reference in tests/test_gpu_loss.py (w1_align_*) and
tests/test_gpu_multirank.py (w2_align_*)."""
import numpy as np
import pytest
import torch

from mmlearn_b200.alignment import _padded_width, alignment_targets, band_ends, modality_alignment_loss


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


@pytest.mark.parametrize("sizes", [[4, 4], [1, 1], [5, 3], [3, 5], [6, 4, 3], [2, 7, 4], [3, 3, 3, 3],
                                   [1, 6, 2], [10, 7, 5], [0, 4], [4, 0, 3]])
def test_band_is_the_reference_target(sizes):
    """t_ij = [i <= j < band_end[i]] is exactly the reference's dense target."""
    n = sum(sizes)
    end = np.asarray(band_ends(sizes))
    j = np.arange(n)
    dense = ((j[None, :] >= j[:, None]) & (j[None, :] < end[:, None])).astype(np.float32)
    assert np.array_equal(dense, _literal_targets(sizes))


def test_padded_width_fits_the_pair_kernel():
    for d in (8, 32, 64, 128, 200, 512, 520, 640, 768, 1024, 1100, 2048):
        dp = _padded_width(d)
        ndc = (dp + 511) // 512
        assert dp >= d and dp % (128 * ndc) == 0 and dp - d < 256


def test_no_cpu_fallback():
    e = {"rgb_embedding": torch.randn(4, 8), "text_embedding": torch.randn(4, 8)}
    with pytest.raises(ValueError, match="no CPU fallback"):
        modality_alignment_loss(e, torch.tensor(10.0), True, False, False)
