#!/usr/bin/env python
"""Headline benchmark: contrastive loss fwd+bwd samples/s at global batch 32768, d=512 (BASELINE.json
config 2), through the drop-in module, on N GPUs of one node.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the reference's CPU arithmetic (oracle port)

Prints ONE JSON line (rank 0).  A "step" = one forward + backward of the loss on one batch of
synthetic embeddings: id matching, L2 normalise + bf16 pack, all-gather,
two fused similarity/LSE passes, two fused recompute-backward passes, gradient unpack.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GLOBAL = 32768
DIM = 512
PEAKS_FALLBACK = {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(PEAKS_FALLBACK)
    d["_source"] = "fallback"
    return d


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def mark(self):
        """Start of the timed region: earlier samples (warm-up) are kept only as a fallback."""
        self.first = len(self.rows)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        first = getattr(self, "first", 0)
        rows = self.rows[first:] if len(self.rows) - first >= 2 else self.rows
        self.rows = rows
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names)
                   if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        # median of the samples taken under load (upper half: idle samples sit at the low clock)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons}


def reference_arm(args):
    """--impl reference: the reference's own CPU arithmetic (pinned oracle port), all host
    threads, on a bounded sample of the same workload (one rank's share: b rows x all N keys)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle.cpu_reference import time_rank_share

    b = args.cpu_rows
    r = time_rank_share(args.n, args.d, b, steps=max(1, args.steps), warmup=min(args.warmup, 1),
                        threads=os.cpu_count())
    line = {
        "impl": "reference", "metric": "contrastive loss fwd+bwd samples/sec @B=32k,d=512",
        "value": r["samples_per_s"], "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": min(args.warmup, 1), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"ContrastiveLoss fwd+bwd, global batch {args.n}, d={args.d}, 1 pair, "
                               "l2_normalize+local_loss+gather_with_grad (BASELINE config 2)"},
        "cpu_baseline": {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"],
                         "kind": "port",
                         "sample": f"one rank's share: {b} query rows x all {args.n} keys, both "
                                   f"directions, fwd+bwd, torch {torch.__version__} CPU fp32"},
        "e2e": {"value": r["samples_per_s"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    # (full-length names: torchrun's own parser treats unknown prefixes such as --n as ambiguous)
    ap.add_argument("--global-batch", dest="n", type=int, default=N_GLOBAL)
    ap.add_argument("--embed-dim", dest="d", type=int, default=DIM)
    ap.add_argument("--cpu-rows", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from mmlearn_b200 import ContrastiveLoss, LossPairSpec, _lib, get_handle
    from mmlearn_b200._lib import check, ptr

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # This program's stdout is the one JSON line: anything libraries print on fd 1 while the
    # bench runs (NCCL's version banner, for one) goes to stderr instead.
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    warmup = max(args.warmup, 3)
    N, d = args.n, args.d
    assert N % world == 0
    b = N // world
    peaks = load_peaks()
    lib = _lib.load()

    # ---- synthetic inputs, generated on the host, pinned ----------------
    g = torch.Generator().manual_seed(1234)
    xa = torch.randn(N, d, generator=g)
    g = torch.Generator().manual_seed(1235)
    xb = 0.5 * xa + torch.randn(N, d, generator=g)
    sl = slice(rank * b, (rank + 1) * b)
    h_a = xa[sl].to(torch.bfloat16).contiguous().pin_memory()
    h_b = xb[sl].to(torch.bfloat16).contiguous().pin_memory()
    h_ids = torch.stack([torch.zeros(b, dtype=torch.long), torch.arange(rank * b, (rank + 1) * b)],
                        1).contiguous().pin_memory()
    h_log_s = torch.tensor(math.log(1 / 0.07)).pin_memory()
    del xa, xb

    fn = ContrastiveLoss(l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=True)
    pairs = [LossPairSpec(("rgb", "text"))]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def one_step(a, bt, ids, log_s):
        a.grad = bt.grad = log_s.grad = None
        loss = fn({"rgb_embedding": a, "text_embedding": bt}, {"rgb": ids, "text": ids},
                  log_s.exp(), pairs)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) device-resident timing: inputs already in HBM ------------------------------------
    d_a = h_a.to(dev).requires_grad_(True)
    d_b = h_b.to(dev).requires_grad_(True)
    d_ids = h_ids.to(dev)
    d_log_s = h_log_s.to(dev).requires_grad_(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before warm-up: nvidia-smi needs ~100 ms to deliver its first sample
    for _ in range(warmup):
        one_step(d_a, d_b, d_ids, d_log_s)
    barrier()
    sampler.mark()
    l0 = lib.mmcl_kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(args.steps)]
    # No host synchronisation between the timed iterations (the bracket is around all K of
    # them): the host runs ahead exactly as it does inside a training step.  The L2 flush is a
    # kernel on the same stream, outside each iteration's event pair.
    span = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    span[0].record()
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # evict L2 between timed iterations (outside the events)
        ev[i][0].record()
        loss = one_step(d_a, d_b, d_ids, d_log_s)
        ev[i][1].record()
    span[1].record()
    barrier()
    launches = (lib.mmcl_kernel_launches() - l0) // args.steps
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    t = torch.tensor([ms, span[0].elapsed_time(span[1])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t[0]) / args.steps
    ms_per_step_with_flush = float(t[1]) / args.steps
    value = N / (ms_per_step * 1e-3)
    loss_value = float(loss.detach())

    # ---- (2) end to end: pinned host buffers -> H2D -> step -> D2H of loss + d(scale) -----------
    # Every step copies its inputs from pinned host memory and the host reads the loss back
    # (and waits for it) before it starts the next step.  Like an input pipeline, the copy of
    # step i+1 is issued on a copy stream before the host waits for step i, into the second of
    # two device buffer sets, so it overlaps step i's kernels.  ONE event pair spans all K steps
    # (copies, L2 flushes, kernels, read-backs).
    e2e_steps = args.steps
    out_host = torch.empty(2, dtype=torch.float32).pin_memory()
    main = torch.cuda.current_stream(dev)
    cs = torch.cuda.Stream(dev)
    sets = [dict(a=torch.empty_like(d_a).requires_grad_(True),
                 b=torch.empty_like(d_b).requires_grad_(True), ids=torch.empty_like(d_ids),
                 log_s=torch.empty_like(d_log_s).requires_grad_(True),
                 ready=torch.cuda.Event(), done=None) for _ in range(2)]
    span2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
    total = e2e_steps + 2  # two untimed steps first

    def issue_copy(i):
        s_ = sets[i % 2]
        with torch.cuda.stream(cs), torch.no_grad():
            if s_["done"] is not None:
                cs.wait_event(s_["done"])  # step i-2 has finished with this buffer set
            if i == 2:
                span2[0].record(cs)  # the timed region starts with the first timed step's copy
            s_["a"].copy_(h_a, non_blocking=True)
            s_["b"].copy_(h_b, non_blocking=True)
            s_["ids"].copy_(h_ids, non_blocking=True)
            s_["log_s"].copy_(h_log_s, non_blocking=True)
            s_["ready"].record(cs)

    barrier()
    issue_copy(0)
    for i in range(total):
        if i + 1 < total:
            issue_copy(i + 1)
        s_ = sets[i % 2]
        flush.fill_(i & 0xFF)
        main.wait_event(s_["ready"])
        loss = one_step(s_["a"], s_["b"], s_["ids"], s_["log_s"])
        out_host.copy_(torch.stack([loss.detach(), s_["log_s"].grad]), non_blocking=True)
        s_["done"] = torch.cuda.Event()
        s_["done"].record(main)
        if i == total - 1:
            span2[1].record(main)
        # the user reads the loss on the host every step (logging): wait for the D2H copy
        main.synchronize()
        e2e_loss = float(out_host[0])
    barrier()
    t2 = torch.tensor([span2[0].elapsed_time(span2[1])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = N / (float(t2) / e2e_steps * 1e-3)
    h2d = h_a.numel() * 2 + h_b.numel() * 2 + h_ids.numel() * 8 + 4
    d2h = 8

    # ---- (3) dominant kernel alone (recompute backward), CUDA events on its stream ------------
    roof = None
    if rank == 0:
        hnd = get_handle(dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        x = torch.nn.functional.normalize(d_a.detach().float(), dim=-1).to(torch.bfloat16)
        y = torch.nn.functional.normalize(
            torch.randn(N, d, device=dev, generator=torch.Generator(device=dev).manual_seed(7)),
            dim=-1).to(torch.bfloat16)
        sc = torch.tensor([1 / 0.07], device=dev)
        xinv = torch.ones(b, device=dev)
        yinv = torch.ones(N, device=dev)
        off_r = torch.full((b,), 30.0, device=dev)
        off_c = torch.full((N,), 30.0, device=dev)
        dx = torch.zeros(b, d, device=dev)
        reps = max(5, min(args.steps, 20))
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(reps)]
        for i in range(reps + 2):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            if i >= 2:
                kev[i - 2][0].record()
            check(lib.mmcl_sim_grad_bwd(hnd, ptr(x), b, ptr(y), N, d, ptr(sc), ptr(xinv), ptr(yinv),
                                        ptr(off_r), ptr(off_c), ptr(dx), st), "bwd")
            if i >= 2:
                kev[i - 2][1].record()
        torch.cuda.synchronize()
        kms = sum(e0.elapsed_time(e1) for e0, e1 in kev) / reps
        flops = 4.0 * b * N * d  # two of the reference's four backward GEMMs per launch
        achieved = flops / (kms * 1e-3) / 1e12
        peak = float(peaks["bf16_tflops"])
        # traffic: dram__bytes_read.sum + dram__bytes_write.sum of this kernel from the committed
        # `ncu --set full` capture at this shape (profiles/r01_ncu_summary.md); algorithmic
        # operand + result bytes are (b + N) * d * 2 + b * d * 4
        traffic = 220.6e6 if (N, d, world) == (32768, 512, 1) else None
        roof = {"bound": "tensor", "kernel": "sim_grad_bwd_pair_kernel", "achieved": achieved,
                "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes": (b + N) * d * 2 + b * d * 4,
                "ms_per_launch": kms, "peak_source": f"{peaks['_source']} burst cuBLAS bf16",
                "step_tflops": 12.0 * b * N * d / (ms_per_step * 1e-3) / 1e12,
                "step_frac_of_sustained": 12.0 * b * N * d / (ms_per_step * 1e-3) / 1e12
                / float(peaks["bf16_tflops_sustained"])}

    # ---- (4) CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_reference import time_rank_share

        r = time_rank_share(N, d, args.cpu_rows, steps=3, warmup=1, threads=os.cpu_count())
        cpu = {"value": r["samples_per_s"], "unit": "samples/s", "cores": r["threads"],
               "kind": "port",
               "sample": f"one rank's share: {args.cpu_rows} query rows x all {N} keys, both "
                         f"directions, fwd+bwd, 3 steps, torch CPU fp32 ({r['ms_per_step']:.0f} ms/step)"}

    # ---- (5) the library baseline on the same GPU: the reference's torch ops (cuBLAS + ATen) -----
    lib_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_reference import time_rank_share

        del flush
        torch.cuda.empty_cache()
        try:
            r = time_rank_share(N, d, N, steps=5, warmup=2, device=str(dev), autocast_bf16=True)
            lib_gpu = {"value": r["samples_per_s"], "unit": "samples/s", "ms_per_step": r["ms_per_step"],
                       "kind": "port of the reference's torch ops on cuda:0 under bf16 autocast "
                               "(cuBLAS GEMMs + ATen cross_entropy; logits materialised in HBM); "
                               "id matching excluded (the reference's needs a 2.1 GB N x N temporary)"}
        except Exception as exc:  # noqa: BLE001 - a baseline must never break the bench line
            lib_gpu = {"value": None, "error": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": "contrastive loss fwd+bwd samples/sec @B=32k,d=512",
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"ContrastiveLoss fwd+bwd, global batch {N}, d={d}, 1 pair, "
                                   "l2_normalize+local_loss+gather_with_grad (BASELINE config 2)",
                       "rows_per_gpu": b, "l2": "flushed between timed iterations (256 MiB fill kernel on the stream, "
                             "outside each iteration's event pair; no host sync between iterations)",
                       "ms_per_step_incl_flush": ms_per_step_with_flush,
                       "parallelism": f"dp{world}"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "how": "one event pair over all K steps incl. H2D copies, L2 flushes and the "
                           "per-step host wait for the loss; the next step's H2D is prefetched on "
                           "a copy stream (double-buffered inputs)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
            "torch_gpu_baseline": lib_gpu,
            "loss": loss_value,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
