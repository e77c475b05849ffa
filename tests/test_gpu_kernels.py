"""GPU parity tests of the individual C-ABI entry points against plain torch fp32 on the same
bf16-rounded inputs, and of the id matcher against the oracle (bit-exact)."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

LOG2E = 1.4426950408889634


def _setup():
    from mmlearn_b200 import _lib, get_handle

    dev = torch.device("cuda:0")
    return _lib.load(), get_handle(dev), dev, torch.cuda.current_stream(dev).cuda_stream


def _rand(n, d, dev, seed, scale=1.0, normalize=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, d, generator=g) * scale
    if normalize:
        x = torch.nn.functional.normalize(x, dim=-1)
    return x.to(dev).to(torch.bfloat16)


@pytest.mark.parametrize("nq,nk,d,s", [
    (256, 256, 512, 14.285), (130, 1000, 64, 5.0), (128, 256, 512, 100.0), (4096, 4096, 512, 14.285),
    (300, 77, 768, 20.0), (512, 2048, 1024, 14.285), (1, 1, 8, 3.0), (777, 3333, 200, -7.5),
    (2048, 40000, 512, 50.0),
])
def test_sim_lse_fwd(nq, nk, d, s):
    from mmlearn_b200._lib import check, ptr

    lib, h, dev, st = _setup()
    q, k = _rand(nq, d, dev, 1), _rand(nk, d, dev, 2)
    scale = torch.tensor([s], device=dev)
    lse = torch.empty(nq, device=dev)
    es = torch.empty(nq, device=dev)
    wsb = lib.mmcl_sim_lse_workspace_bytes(h, nq, nk, d)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    # un-normalised rows + fp32 inverse norms applied in the epilogue (the production path)
    qn = (torch.rand(nq, device=dev) * 3 + 0.5)
    kn = (torch.rand(nk, device=dev) * 3 + 0.5)
    q = (q.float() * qn[:, None]).to(torch.bfloat16)
    k = (k.float() * kn[:, None]).to(torch.bfloat16)
    qinv = (1 / q.float().norm(dim=1)).contiguous()
    kinv = (1 / k.float().norm(dim=1)).contiguous()
    check(lib.mmcl_sim_lse_fwd(h, ptr(q), nq, ptr(k), nk, d, ptr(scale), ptr(qinv), ptr(kinv),
                               ptr(lse), ptr(es), ptr(ws), wsb, st), "fwd")
    torch.cuda.synchronize()
    S = s * ((q.float() * qinv[:, None]) @ (k.float() * kinv[:, None]).T)
    ref_lse = torch.logsumexp(S, dim=1)
    ref_es = (torch.softmax(S, dim=1) * S).sum(1)
    # tolerance: fp32 accumulation + ex2.approx; logits up to |s|
    assert torch.allclose(lse, ref_lse, rtol=1e-4, atol=2e-4), (lse - ref_lse).abs().max()
    assert torch.allclose(es, ref_es, rtol=2e-3, atol=2e-3), (es - ref_es).abs().max()


@pytest.mark.parametrize("nx,ny,d,s,mode", [
    (128, 256, 256, 10.0, 3), (256, 512, 512, 14.285, 3), (130, 1000, 64, 5.0, 3),
    (300, 77, 768, 20.0, 1), (512, 2048, 1024, 14.285, 2), (4096, 4096, 512, 14.285, 3),
    (1000, 3000, 384, 30.0, 3), (64, 64, 8, 2.0, 3),
])
def test_sim_grad_bwd(nx, ny, d, s, mode):
    from mmlearn_b200._lib import check, ptr

    lib, h, dev, st = _setup()
    x, y = _rand(nx, d, dev, 3), _rand(ny, d, dev, 4)
    scale = torch.tensor([s], device=dev)
    x = (x.float() * (torch.rand(nx, device=dev) * 3 + 0.5)[:, None]).to(torch.bfloat16)
    y = (y.float() * (torch.rand(ny, device=dev) * 3 + 0.5)[:, None]).to(torch.bfloat16)
    xinv = (1 / x.float().norm(dim=1)).contiguous()
    yinv = (1 / y.float().norm(dim=1)).contiguous()
    yh = y.float() * yinv[:, None]
    S = s * ((x.float() * xinv[:, None]) @ yh.T)
    lse_r = torch.logsumexp(S, 1)
    lse_c = torch.logsumexp(S, 0)
    off_r = (lse_r * LOG2E + np.log2(2.0 * nx)).contiguous()
    off_c = (lse_c * LOG2E + np.log2(2.0 * ny)).contiguous()
    dx = torch.zeros(nx, d, device=dev)
    check(lib.mmcl_sim_grad_bwd(h, ptr(x), nx, ptr(y), ny, d, ptr(scale), ptr(xinv), ptr(yinv),
                                ptr(off_r) if mode & 1 else None, ptr(off_c) if mode & 2 else None,
                                ptr(dx), st), "bwd")
    torch.cuda.synchronize()
    G = torch.zeros_like(S)
    if mode & 1:
        G += torch.exp(S - lse_r[:, None]) / (2.0 * nx)
    if mode & 2:
        G += torch.exp(S - lse_c[None, :]) / (2.0 * ny)
    ref = G @ yh
    err = (dx - ref).norm() / ref.norm()
    # G is rounded to bf16 (2^-9 relative per element) before the second MMA
    assert err < 4e-3, float(err)
    cos = torch.nn.functional.cosine_similarity(dx.flatten(), ref.flatten(), dim=0)
    assert cos > 0.9999


def test_sim_grad_bwd_wide_offset_range():
    """Offsets spread over more than 2*40 binades (duplicated rows at s=100): the kernel must
    leave its one-exponential fast form and use the exact two-exponential form."""
    from mmlearn_b200._lib import check, ptr

    lib, h, dev, st = _setup()
    nx, ny, d, s = 256, 512, 128, 100.0
    x, y = _rand(nx, d, dev, 5), _rand(ny, d, dev, 6)
    y[:64] = x[:64]
    scale = torch.tensor([s], device=dev)
    xinv = (1 / x.float().norm(dim=1)).contiguous()
    yinv = (1 / y.float().norm(dim=1)).contiguous()
    yh = y.float() * yinv[:, None]
    S = s * ((x.float() * xinv[:, None]) @ yh.T)
    lse_r, lse_c = torch.logsumexp(S, 1), torch.logsumexp(S, 0)
    off_r = (lse_r * LOG2E + np.log2(2.0 * nx)).contiguous()
    off_c = (lse_c * LOG2E + np.log2(2.0 * ny)).contiguous()
    assert float(max(off_r.max(), off_c.max()) - min(off_r.min(), off_c.min())) > 80.0
    dx = torch.zeros(nx, d, device=dev)
    check(lib.mmcl_sim_grad_bwd(h, ptr(x), nx, ptr(y), ny, d, ptr(scale), ptr(xinv), ptr(yinv),
                                ptr(off_r), ptr(off_c), ptr(dx), st), "bwd")
    torch.cuda.synchronize()
    G = torch.exp(S - lse_r[:, None]) / (2.0 * nx) + torch.exp(S - lse_c[None, :]) / (2.0 * ny)
    ref = G @ yh
    err = (dx - ref).norm() / ref.norm()
    assert err < 4e-3, float(err)


def test_match_ids_bit_exact():
    from mmlearn_b200 import find_matching_indices
    from oracle.contrastive_oracle import find_matching_indices_np, find_matching_indices_sorted
    from tests.golden_util import GOLDEN

    dev = torch.device("cuda:0")
    # reference known-answer vectors: /root/reference/tests/datasets/test_example.py:147-178
    a = torch.tensor([(0, 0), (0, 1), (1, 0), (1, 1)], device=dev)
    b = torch.tensor([(1, 0), (1, 1), (2, 0), (2, 1), (2, 2)], device=dev)
    ia, ib = find_matching_indices(a, b)
    assert ia.tolist() == [2, 3] and ib.tolist() == [0, 1]
    a2 = torch.tensor([(0, 0), (0, 1), (1, 0), (1, 1), (1, 1)], device=dev)
    ia, ib = find_matching_indices(a2, b)
    assert ia.tolist() == [2, 3, 4] and ib.tolist() == [0, 1, 1]
    ia, ib = find_matching_indices(a, torch.tensor([(2, 0), (2, 1), (2, 2)], device=dev))
    assert ia.numel() == 0 and ib.numel() == 0 and ia.dtype == torch.long
    with pytest.raises(TypeError, match="Expected inputs to be tensors, but got.*"):
        find_matching_indices([(0, 0)], b)
    with pytest.raises(ValueError, match=r"Expected argument `first_example_ids` to be a tensor of shape \(N, 2\).*"):
        find_matching_indices(torch.tensor([(0,), (0,)], device=dev), b)
    # golden outputs of the reference function
    z = np.load(f"{GOLDEN}/match_random.npz")
    for i in range(len([k for k in z.files if k.endswith("_ia")])):
        ia, ib = find_matching_indices(torch.from_numpy(z[f"c{i}_a"]).to(dev),
                                       torch.from_numpy(z[f"c{i}_b"]).to(dev))
        assert np.array_equal(ia.cpu().numpy(), z[f"c{i}_ia"])
        assert np.array_equal(ib.cpu().numpy(), z[f"c{i}_ib"])
    # random: duplicates (cross products), misses, int64 range, permutations, large sizes
    rng = np.random.RandomState(0)
    for n, m, hi in [(1000, 900, 300), (5000, 5000, 10 ** 6), (32768, 32768, 2 ** 62), (1, 1, 1),
                     (0, 5, 3), (300, 300, 5)]:
        a = np.stack([rng.randint(0, 3, n), rng.randint(0, hi, n, dtype=np.int64)], 1).astype(np.int64)
        b = np.stack([rng.randint(0, 3, m), rng.randint(0, hi, m, dtype=np.int64)], 1).astype(np.int64)
        if n == m and n > 100:
            b[: n // 2] = a[rng.permutation(n)[: n // 2]]
        fn = find_matching_indices_np if n * m < 4e6 else find_matching_indices_sorted
        ra, rb = fn(a, b)
        ia, ib = find_matching_indices(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
        assert np.array_equal(ia.cpu().numpy(), ra) and np.array_equal(ib.cpu().numpy(), rb)
    # identity fast-path flag
    from mmlearn_b200.loss import _Match, get_handle

    ids = torch.stack([torch.zeros(4096, dtype=torch.long), torch.randperm(4096)], 1).to(dev)
    mt = _Match(get_handle(dev), ids, ids.clone())
    assert mt.info.tolist() == [4096, 1, 4096, 4096]
    assert _Match(get_handle(dev), ids, ids.clone(), quick=True).info.tolist() == [4096, 1, 4096, 4096]
    ids2 = ids.clone()
    ids2[5] = ids2[6]
    assert _Match(get_handle(dev), ids, ids2).info.tolist()[1] == 0
    assert _Match(get_handle(dev), ids, ids2, quick=True).info.tolist() == [0, 0, 4096, 4096]
    assert _Match(get_handle(dev), ids2, ids2, quick=True).info.tolist() == [0, 0, 4096, 4096]  # duplicate id


@pytest.mark.parametrize("d", [8, 64, 128, 160, 192, 256, 320, 384, 448, 520, 768, 2048])
@pytest.mark.parametrize("src_dtype", [torch.float32, torch.bfloat16])
def test_pack_and_unpack_every_width(d, src_dtype):
    """Row packing (inverse norms) and gradient unpack (positive term + normalise backward) against
    torch fp32 for widths whose last 256-column group is partly filled: the warp reduction of
    both kernels must see every column (d % 256 in {32, 64, 96, ...} exercised here)."""
    from mmlearn_b200._lib import check, ptr

    lib, h, dev, st = _setup()
    n = 301
    g = torch.Generator().manual_seed(d)
    src = (torch.randn(n, d, generator=g) * 1.7).to(src_dtype).to(dev)
    dt = {torch.float32: 0, torch.bfloat16: 1}[src_dtype]
    rows = torch.empty(n, d, dtype=torch.bfloat16, device=dev)
    inv = torch.empty(n, device=dev)
    check(lib.mmcl_pack_rows(h, ptr(src), dt, n, d, None, n, 1, ptr(rows), ptr(inv), st), "pack")
    want_inv = 1 / src.float().norm(dim=1)
    torch.testing.assert_close(inv, want_inv, rtol=1e-5, atol=0)
    assert torch.equal(rows, src.to(torch.bfloat16))
    dacc = (torch.randn(n, d, generator=g) * 1e-3).to(dev)
    partner = torch.randn(n, d, generator=g).to(torch.bfloat16).to(dev)
    pinv = torch.rand(n, generator=g).to(dev) + 0.5
    scale = torch.tensor([14.285], device=dev)
    gout = torch.tensor(0.75, device=dev)
    out = torch.empty(n, d, dtype=src_dtype, device=dev)
    weight, pos_coef = 1.3, 2e-4
    check(lib.mmcl_unpack_grad(h, ptr(dacc), ptr(partner), ptr(pinv), n, d, weight, pos_coef, ptr(gout),
                               ptr(scale), ptr(src), dt, ptr(inv), 1, None, ptr(out), dt, 1, st), "unpack")
    torch.cuda.synchronize()
    gh = weight * 0.75 * 14.285 * (dacc - pos_coef * pinv[:, None] * partner.float())
    xh = src.float() * want_inv[:, None]
    want = want_inv[:, None] * (gh - xh * (xh * gh).sum(1, keepdim=True))
    tol = 1e-5 if src_dtype == torch.float32 else 1e-2
    assert float((out.float() - want).norm() / want.norm()) < tol
