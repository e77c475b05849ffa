"""All-gather transports of the loss, timed on N GPUs (torchrun): every rank contributes `rows` bf16
rows of `d` (4 MB at rows = 4096, d = 512 -- one modality of config 2 at 8 GPUs).

  a) pack_rows_symm : normalise + bf16 + store to every peer's arena buffer, then one fence
  b) symm_put       : copy of already packed rows to every peer, then one fence
  c) NCCL           : mmcl_all_gather_rows (ncclAllGather) on the handle's communicator

Each variant is timed with CUDA events from the first launch to the end of the fence / collective,
max over ranks, median of `reps`.
usage: torchrun --nproc-per-node N tools/push_bench.py [rows] [d] [reps]
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from mmlearn_b200 import _lib, get_handle
from mmlearn_b200._lib import check, counts_array, ptr

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
d = int(sys.argv[2]) if len(sys.argv) > 2 else 512
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 30
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
assert world > 1, "run under torchrun with at least 2 ranks"
torch.cuda.set_device(local_rank)
dev = torch.device("cuda", local_rank)
dist.init_process_group("nccl", device_id=dev)
lib, h = _lib.load(), get_handle(dev)
st = torch.cuda.current_stream(dev).cuda_stream
N = rows * world
need = N * d * 2 + N * 4 + (1 << 20)
check(lib.mmcl_symm_reserve(h, need, st), "reserve")
src = torch.randn(rows, d, device=dev).to(torch.bfloat16)
packed = src.clone()
full = torch.empty(N, d, dtype=torch.bfloat16, device=dev)
off_rows, off_inv = 0, N * d * 2


def timed(fn, label):
    ts = []
    for i in range(reps + 3):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i & 1)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    t = torch.tensor(sorted(ts)[len(ts) // 2], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        mb = rows * d * 2 * (world - 1) / 1e6
        print(f"{label:34s} {float(t) * 1e3:8.1f} us   ({mb:.1f} MB out per rank, "
              f"{mb / 1e3 / (float(t) * 1e-3):.0f} GB/s egress)")


def a(parity):
    check(lib.mmcl_pack_rows_symm(h, parity, ptr(src), 1, rows, d, None, rows, 1,
                                  off_rows + rank * rows * d * 2, off_inv + rank * rows * 4, st), "pack")
    check(lib.mmcl_symm_fence(h, st), "fence")


def b(parity):
    check(lib.mmcl_symm_put(h, parity, ptr(packed), rows * d * 2, off_rows + rank * rows * d * 2, st), "put")
    check(lib.mmcl_symm_fence(h, st), "fence")


def c(parity):
    full[rank * rows:(rank + 1) * rows].copy_(packed)
    check(lib.mmcl_all_gather_rows(h, full.data_ptr() + rank * rows * d * 2, ptr(full),
                                   counts_array([rows] * world), d * 2, st), "ag")


def f(parity):
    check(lib.mmcl_symm_fence(h, st), "fence")


def push(mode, blocks=0, threads=0):
    def run(parity):
        check(lib.mmcl_debug_push(h, parity, off_rows + rank * rows * d * 2, rows * d * 2, mode, blocks,
                                  threads, st), "push")
        check(lib.mmcl_symm_fence(h, st), "fence")
    return run


timed(f, "fence alone")
for blocks, threads in ((148, 64), (148, 128), (148, 256), (296, 256), (592, 256), (1184, 256)):
    timed(push(0, blocks, threads), f"store kernel {blocks}x{threads} + fence")
timed(push(1), "cudaMemcpyAsync per peer, 1 stream")
timed(push(2), "cudaMemcpyAsync per peer, side streams")
timed(a, "pack_rows_symm + fence")
timed(b, "symm_put + fence")
timed(c, "NCCL all-gather (in place)")
dist.barrier()
dist.destroy_process_group()
