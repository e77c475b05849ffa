"""Kernel-level timings through the C ABI (CUDA events, one GPU): the two hot kernels at one
rank's share of config 2 for W = 1 and W = 8, the identity check and the pack kernel.
usage: python tools/kernel_bench.py [reps]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mmlearn_b200 import _lib, get_handle
from mmlearn_b200._lib import check, ptr

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
lib, h, st = _lib.load(), get_handle(dev), torch.cuda.current_stream(dev).cuda_stream
d, N = 512, 32768
g = torch.Generator(device=dev).manual_seed(1)
y = torch.randn(N, d, device=dev, generator=g).to(torch.bfloat16)
yinv = (1 / y.float().norm(dim=1)).contiguous()
scale = torch.tensor([1 / 0.07], device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, label, flops=None):
    for _ in range(3):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        flush.fill_(1)
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in ev)
    med = ts[len(ts) // 2]
    extra = f"  {flops / med / 1e9:8.1f} TFLOP/s" if flops else ""
    print(f"{label:46s} median {med * 1e3:9.1f} us  min {ts[0] * 1e3:9.1f} us{extra}")


for b in (32768, 4096, 16384):
    x = torch.randn(b, d, device=dev, generator=g).to(torch.bfloat16)
    xinv = (1 / x.float().norm(dim=1)).contiguous()
    lse = torch.empty(b, device=dev)
    es = torch.empty(b, device=dev)
    wsb = lib.mmcl_sim_lse_workspace_bytes(h, b, N, d)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    timeit(lambda: check(lib.mmcl_sim_lse_fwd(h, ptr(x), b, ptr(y), N, d, ptr(scale), ptr(xinv),
                                              ptr(yinv), ptr(lse), ptr(es), ptr(ws), wsb, st), "f"),
           f"sim_lse_fwd      nq={b} nk={N} d={d}", 2.0 * b * N * d)
    lse_c = torch.full((N,), math.log(N) + 2.0, device=dev)
    dx = torch.empty(b, d, device=dev)
    timeit(lambda: check(lib.mmcl_sim_grad_bwd_lse(h, ptr(x), b, ptr(y), N, d, ptr(scale), ptr(xinv),
                                                   ptr(yinv), ptr(lse), 2.0 * b, ptr(lse_c), 2.0 * b,
                                                   ptr(dx), 1, st), "b"),
           f"sim_grad_bwd_lse (memset + kernel) nx={b}", 4.0 * b * N * d)
    thr = torch.zeros(b, device=dev)
    rank = torch.empty(b, dtype=torch.long, device=dev)
    timeit(lambda: check(lib.mmcl_sim_rank_fwd(h, ptr(x), b, ptr(y), N, d, ptr(xinv), ptr(yinv),
                                               ptr(thr), None, ptr(rank), ptr(ws), wsb, st), "r"),
           f"sim_rank_fwd     nq={b} nk={N} d={d}", 2.0 * b * N * d)
    ids = torch.stack([torch.zeros(b, dtype=torch.long), torch.arange(b)], 1).to(dev)
    info = torch.empty(4, dtype=torch.long, device=dev)
    iw = lib.mmcl_identity_check_workspace_bytes(b)
    iws = torch.empty(iw, dtype=torch.uint8, device=dev)
    timeit(lambda: check(lib.mmcl_identity_check(h, ptr(ids), b, ptr(ids), b, ptr(info), ptr(iws), iw, st), "i"),
           f"identity_check   n={b}")
    src = torch.randn(b, d, device=dev).to(torch.bfloat16)
    dst = torch.empty(b, d, dtype=torch.bfloat16, device=dev)
    inv = torch.empty(b, device=dev)
    timeit(lambda: check(lib.mmcl_pack_rows(h, ptr(src), 1, b, d, None, b, 1, ptr(dst), ptr(inv), st), "p"),
           f"pack_rows        n={b} (bf16 in)")
    gout = torch.ones((), device=dev)
    grad = torch.empty(b, d, dtype=torch.bfloat16, device=dev)
    timeit(lambda: check(lib.mmcl_unpack_grad(h, ptr(dx), ptr(x), ptr(xinv), b, d, 1.0, 1e-4, ptr(gout),
                                              ptr(scale), ptr(src), 1, ptr(inv), 1, None, ptr(grad), 1, 1,
                                              st), "u"),
           f"unpack_grad      n={b} (bf16 out, direct)")

# kernel durations without launch gaps (CUPTI), for the small shape where gaps matter
if os.environ.get("MMCL_KB_PROFILE", "1") == "1":
    from torch.profiler import ProfilerActivity, profile

    b = 4096
    x = torch.randn(b, d, device=dev, generator=g).to(torch.bfloat16)
    xinv = (1 / x.float().norm(dim=1)).contiguous()
    lse = torch.empty(b, device=dev)
    es = torch.empty(b, device=dev)
    wsb = lib.mmcl_sim_lse_workspace_bytes(h, b, N, d)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    lse_c = torch.full((N,), math.log(N) + 2.0, device=dev)
    dx = torch.empty(b, d, device=dev)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            flush.fill_(1)
            check(lib.mmcl_sim_lse_fwd(h, ptr(x), b, ptr(y), N, d, ptr(scale), ptr(xinv), ptr(yinv),
                                       ptr(lse), ptr(es), ptr(ws), wsb, st), "f")
            flush.fill_(1)
            check(lib.mmcl_sim_grad_bwd_lse(h, ptr(x), b, ptr(y), N, d, ptr(scale), ptr(xinv), ptr(yinv),
                                            ptr(lse), 2.0 * b, ptr(lse_c), 2.0 * b, ptr(dx), 1, st), "b")
        torch.cuda.synchronize()
    evs = sorted((e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA),
                 key=lambda e: e.time_range.start)
    prev = evs[0].time_range.start
    for e in evs:
        print(f"  dur {e.time_range.end - e.time_range.start:8.1f}  gap {e.time_range.start - prev:7.1f}  {e.name[:60]}")
        prev = e.time_range.end
