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
import gc
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
# BASELINE.json configs 2-4: modalities, pairs (a, b, weight), logit scales
CONFIGS = {
    2: dict(n=32768, d=512, mods=["rgb", "text"], pairs=[("rgb", "text", 1.0)], scales=[1 / 0.07]),
    3: dict(n=131072, d=1024, mods=["rgb", "text"], pairs=[("rgb", "text", 1.0)],
            scales=[1 / 0.07]),
    4: dict(n=16384, d=768, mods=["rgb", "dna", "text"],
            pairs=[("rgb", "dna", 1.0), ("rgb", "text", 0.5), ("dna", "text", 2.0)],
            scales=[1 / 0.07, 10.0, 50.0]),
}
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


def workload_name(cfg_id, N, d, P, world, weak):
    base = {2: "BASELINE config 2", 3: "BASELINE config 3", 4: "BASELINE config 4"}[cfg_id]
    w = (f"ContrastiveLoss fwd+bwd, global batch {N}, d={d}, {P} pair{'s' if P > 1 else ''}, "
         f"l2_normalize+local_loss+gather_with_grad ({base})")
    if weak:
        w += f"; weak-scaling series: {N // world} rows per GPU fixed"
    return w


def metric_name(cfg_id, N, d):
    if cfg_id == 2 and N == N_GLOBAL and d == DIM:
        return "contrastive loss fwd+bwd samples/sec @B=32k,d=512"
    return f"contrastive loss fwd+bwd samples/sec @B={N},d={d} (config {cfg_id})"


def resolve(args):
    """(N, d, modalities, pairs, scales, weak) for this launch."""
    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N = args.n if args.n else cfg["n"]
    d = args.d if args.d else cfg["d"]
    weak = args.rows_per_gpu > 0
    if weak:
        N = args.rows_per_gpu * world
    return N, d, cfg["mods"], cfg["pairs"], cfg["scales"], weak


def reference_arm(args):
    """--impl reference: the reference's own CPU arithmetic (pinned oracle port), all host
    threads, on a bounded sample of the same workload (one rank's share: b rows x all N keys)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch

    from oracle.cpu_reference import time_rank_share

    N, d, mods, pairs, scales, weak = resolve(args)
    b = args.cpu_rows
    warm = max(args.warmup, 3)  # same warm-up rule as the repo arm
    r = time_rank_share(N, d, b, steps=max(1, args.steps), warmup=warm, threads=os.cpu_count())
    P = len(pairs)
    line = {
        "impl": "reference", "metric": metric_name(args.config, N, d),
        "value": r["samples_per_s"] / P, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": r["ms_per_step"] * P,
        "higher_is_better": True, "scaling": "weak" if weak else "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args.config, N, d, P, args.gpus, weak)},
        "cpu_baseline": {"value": r["samples_per_s"] / P, "unit": "samples/s",
                         "cores": r["threads"], "kind": "port",
                         "sample": f"one rank's share: {b} query rows x all {N} keys, both "
                                   f"directions, fwd+bwd, one pair timed"
                                   + (f" and scaled by the {P} pairs" if P > 1 else "")
                                   + f", torch {torch.__version__} CPU fp32"},
        "e2e": {"value": r["samples_per_s"] / P, "unit": "samples/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ncu_traffic(kernel: str, nx: int, ny: int, d: int):
    """dram bytes per launch of `kernel` at this shape from the committed ncu captures
    (profiles/ncu_traffic.json, written by tools/ncu_top.py from `ncu --set full` CSV pages)."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(p):
        return None, None
    try:
        for row in json.load(open(p)):
            if row["kernel"] == kernel and (row["nx"], row["ny"], row["d"]) == (nx, ny, d):
                return row["dram_bytes"], row.get("source")
    except Exception:  # noqa: BLE001
        pass
    return None, None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS),
                    help="BASELINE.json config: 2 (headline), 3 (131072 x 1024), 4 (3 modalities)")
    # (full-length names: torchrun's own parser treats unknown prefixes such as --n as ambiguous)
    ap.add_argument("--global-batch", dest="n", type=int, default=0)
    ap.add_argument("--embed-dim", dest="d", type=int, default=0)
    ap.add_argument("--rows-per-gpu", type=int, default=0,
                    help="weak-scaling series: fixed rows per GPU, global batch = rows x GPUs")
    ap.add_argument("--cpu-rows", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    from mmlearn_b200 import ContrastiveLoss, LossPairSpec, Modalities, _lib, get_handle
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
    N, d, mods, pair_cfg, scales, weak = resolve(args)
    assert N % world == 0
    b = N // world
    P, M = len(pair_cfg), len(mods)
    peaks = load_peaks()
    lib = _lib.load()
    for m in mods:
        if not Modalities.has_modality(m):
            Modalities.register_modality(m)  # "dna", as projects/bioscan_clip/dataset.py:18

    # ---- synthetic inputs, generated on the host, pinned ----------------
    # Every rank generates the whole global batch (the parity check below needs all keys) and
    # keeps its own row block as the step's input.
    sl = slice(rank * b, (rank + 1) * b)
    full_bf16, h_emb = {}, {}
    base = None
    for mi, m in enumerate(mods):
        g = torch.Generator().manual_seed(1234 + mi)
        x = torch.randn(N, d, generator=g)
        if base is None:
            base = x
        else:
            x = 0.5 * base + x
        full_bf16[m] = x.to(torch.bfloat16)
        h_emb[m] = full_bf16[m][sl].contiguous().pin_memory()
    del base, x
    h_ids = torch.stack([torch.zeros(b, dtype=torch.long), torch.arange(rank * b, (rank + 1) * b)],
                        1).contiguous().pin_memory()
    h_log_s = torch.tensor([math.log(s_) for s_ in scales]).pin_memory()
    if P == 1:
        h_log_s = h_log_s.reshape(()).pin_memory()

    fn = ContrastiveLoss(l2_normalize=True, local_loss=True, gather_with_grad=True, cache_labels=True)
    pairs = [LossPairSpec((a_, b_), w_) for a_, b_, w_ in pair_cfg]
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    host_split = {"fwd": [], "bwd": []}  # host seconds spent inside forward / backward calls

    def one_step(embs, ids, log_s):
        for t_ in embs.values():
            t_.grad = None
        log_s.grad = None
        t0_ = time.perf_counter()
        loss = fn({f"{m}_embedding": embs[m] for m in mods}, {m: ids for m in mods},
                  log_s.exp(), pairs)
        t1_ = time.perf_counter()
        loss.backward()
        host_split["fwd"].append(t1_ - t0_)
        host_split["bwd"].append(time.perf_counter() - t1_)
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- (1) device-resident timing: inputs already in HBM ------------------------------------
    d_emb = {m: h_emb[m].to(dev).requires_grad_(True) for m in mods}
    d_ids = h_ids.to(dev)
    d_log_s = h_log_s.to(dev).requires_grad_(True)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before warm-up: nvidia-smi needs ~100 ms to deliver its first sample
    for _ in range(warmup):
        one_step(d_emb, d_ids, d_log_s)
    # A full (generation-2) garbage collection of a process that has imported torch takes tens of
    # milliseconds -- a hundred 1 ms steps.  As training loops that care about step-time jitter
    # do, the collector is run now and kept off inside the timed regions.
    gc.collect()
    gc.disable()
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
        loss = one_step(d_emb, d_ids, d_log_s)
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
    # Every step copies its inputs from pinned host memory and copies the loss and d(log scale)
    # back to pinned host memory.  Like an input pipeline, the H2D copy of step i+1 is issued on a
    # copy stream before step i is enqueued, into the second of two device buffer sets.  The host
    # reads each step's loss ONE STEP LATE (it waits for the D2H copy of step i-1 after it has
    # enqueued step i), the way a training loop logs without draining the device every step; the
    # last step's value is waited for inside the timed region.  ONE event pair spans all K steps
    # (copies, L2 flushes, kernels, read-backs).
    e2e_steps = args.steps
    main = torch.cuda.current_stream(dev)
    cs = torch.cuda.Stream(dev)
    sets = [dict(emb={m: torch.empty_like(d_emb[m]).requires_grad_(True) for m in mods},
                 ids=torch.empty_like(d_ids),
                 log_s=torch.empty_like(d_log_s).requires_grad_(True),
                 ready=torch.cuda.Event(), done=None) for _ in range(2)]
    out_host = [torch.empty(1 + max(P, 1), dtype=torch.float32).pin_memory() for _ in range(2)]

    def run_e2e():
        """One pass of the end-to-end loop; returns (samples/s, host period stats, losses)."""
        span2 = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        total = e2e_steps + 2  # two untimed steps first
        for s_ in sets:
            s_["done"] = None

        def issue_copy(i):
            s_ = sets[i % 2]
            with torch.cuda.stream(cs), torch.no_grad():
                if s_["done"] is not None:
                    cs.wait_event(s_["done"])  # step i-2 has finished with this buffer set
                if i == 2:
                    span2[0].record(cs)  # the timed region starts with the first timed step's copy
                for m in mods:
                    s_["emb"][m].copy_(h_emb[m], non_blocking=True)
                s_["ids"].copy_(h_ids, non_blocking=True)
                s_["log_s"].copy_(h_log_s, non_blocking=True)
                s_["ready"].record(cs)

        barrier()
        issue_copy(0)
        read_ev = [None, None]
        losses, host_t = [], []
        host_split["fwd"].clear()
        host_split["bwd"].clear()
        for i in range(total):
            if i + 1 < total:
                issue_copy(i + 1)
            s_ = sets[i % 2]
            flush.fill_(i & 0xFF)
            main.wait_event(s_["ready"])
            loss_ = one_step(s_["emb"], s_["ids"], s_["log_s"])
            out_host[i % 2].copy_(torch.cat([loss_.detach().reshape(1),
                                             s_["log_s"].grad.reshape(-1)]), non_blocking=True)
            s_["done"] = torch.cuda.Event()
            s_["done"].record(main)
            read_ev[i % 2] = s_["done"]
            if i == total - 1:
                span2[1].record(main)
            if i >= 1:  # the host consumes step i-1's loss now (logging lags by one step)
                read_ev[(i - 1) % 2].synchronize()
                losses.append(float(out_host[(i - 1) % 2][0]))
            host_t.append(time.perf_counter())
        read_ev[(total - 1) % 2].synchronize()
        losses.append(float(out_host[(total - 1) % 2][0]))
        barrier()
        t2 = torch.tensor([span2[0].elapsed_time(span2[1])], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        dt = sorted((b_ - a_) * 1e3 for a_, b_ in zip(host_t[2:-1], host_t[3:]))
        med = lambda v: sorted(v)[len(v) // 2] * 1e3 if v else None  # noqa: E731
        host = {"median": dt[len(dt) // 2] if dt else None, "max": dt[-1] if dt else None,
                "in_forward_call": med(host_split["fwd"][2:]),
                "in_backward_call": med(host_split["bwd"][2:]),
                "note": "rank 0's host loop period; a single long step (allocator, page faults) "
                        "shows up as max >> median"}
        return N / (float(t2) / e2e_steps * 1e-3), host, losses

    e2e_value, e2e_host, e2e_losses = run_e2e()
    # the same loop with the opt-in deferred verification (ContrastiveLoss.deferred_verification):
    # forward() returns without waiting for the device, so the host runs a whole step ahead
    fn.deferred_verification = True
    e2e_deferred, e2e_deferred_host, losses_d = run_e2e()
    fn.deferred_verification = False
    fn._resolve_deferred()
    gc.enable()
    h2d = sum(h_emb[m].numel() * 2 for m in mods) + h_ids.numel() * 8 + 4 * max(P, 1)
    d2h = 4 * (1 + max(P, 1))
    e2e_ok = all(math.isfinite(v) and abs(v - loss_value) <= 1e-3 * abs(loss_value)
                 for v in e2e_losses + losses_d)

    # ---- (3) parity self-check, outside every timed region, on every rank ----------------------
    # One more step through the same module (so: the same steady-state path that was timed),
    # compared with a chunked fp32 torch evaluation of THIS rank's loss, d(scale) and 256 sampled
    # gradient rows per modality (tools/chunked_check.py; reference semantics contrastive.py:196-212,
    # 321-332).  Tolerances are the north-star's: loss 1e-3, gradient norms 1e-2, d(scale) 1e-2
    # relative; labels bit-exact.  The maximum over ranks is reported; a failure exits non-zero.
    parity = None
    if not args.no_parity:
        from tools.chunked_check import compare, rank_check

        loss = one_step(d_emb, d_ids, d_log_s)
        torch.cuda.synchronize()
        # d(loss)/d(scale) from d(loss)/d(log scale): s = exp(log s)
        ds = (d_log_s.grad.reshape(-1) / d_log_s.detach().exp().reshape(-1)).tolist()
        del flush
        torch.cuda.empty_cache()
        full_dev = {m: full_bf16[m].to(dev) for m in mods}
        ref = rank_check(full_dev, [(a_, b_, w_, s_) for (a_, b_, w_), s_ in
                                    zip(pair_cfg, scales * (P if len(scales) == 1 else 1))],
                         rank, world, chunk=4096 if N <= 65536 else 2048)
        cmpd = compare(ref, float(loss.detach()), ds, {m: d_emb[m].grad for m in mods})
        want_labels = torch.arange(N, device=dev)[rank * b:(rank + 1) * b if rank + 1 < world
                                                  else None]
        got_labels = fn._labels.get(dev)
        labels_ok = got_labels is not None and torch.equal(got_labels, want_labels)
        errs = torch.tensor([cmpd["loss_rel"], cmpd["grad_rel"], cmpd["dscale_rel"],
                             1.0 - cmpd["grad_cos_min"], 0.0 if labels_ok else 1.0,
                             0.0 if e2e_ok else 1.0],
                            device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        e_ = errs.tolist()
        parity = {"loss_rel": e_[0], "grad_rel": e_[1], "dscale_rel": e_[2],
                  "grad_cos_min": 1.0 - e_[3], "labels_ok": e_[4] == 0.0,
                  "e2e_losses_match": e_[5] == 0.0,
                  "tolerance": {"loss_rel": 1e-3, "grad_rel": 1e-2, "dscale_rel": 1e-2},
                  "how": "max over all ranks; every rank checks its own loss, d(scale), labels and "
                         "256 sampled gradient rows per modality against a chunked fp32 torch "
                         "evaluation of the same bf16-rounded inputs (tools/chunked_check.py)",
                  "steady_steps": int(fn.steady_steps)}
        parity["ok"] = bool(e_[0] <= 1e-3 and e_[1] <= 1e-2 and e_[2] <= 1e-2 and e_[4] == 0.0
                            and e_[5] == 0.0 and e_[3] < 1e-3)
        del full_dev, ref
        torch.cuda.empty_cache()
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)

    # ---- (4) dominant kernel alone (recompute backward), CUDA events on its stream ------------
    # The launch the step itself issues: one more steady-state forward without the eager backward,
    # then the step's fused backward launch (every (pair, side) problem in one
    # sim_grad_bwd_pair_kernel launch, `mmcl_step_grad_kernels`) timed on its own with CUDA events,
    # L2 flushed before every launch.  It reads this step's real log-sum-exp vectors, so it takes
    # whichever exponential path the data selects.  (With several ranks the call would need the
    # peers in lock step; the roofline entry is a one-GPU measurement.)
    roof = None
    if rank == 0 and world == 1:
        fn.eager_backward = False
        loss = fn({f"{m}_embedding": d_emb[m] for m in mods}, {m: d_ids for m in mods},
                  d_log_s.exp(), pairs)
        run = getattr(loss.grad_fn, "run", None)
        fast = getattr(run, "fast", None)
        kms, sides = None, 2 * P
        if fast is not None:
            sp, buf_parity, es_, ws_ = fast
            st = torch.cuda.current_stream(dev).cuda_stream
            reps = max(5, min(args.steps, 20))
            kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                   for _ in range(reps)]
            for i in range(reps + 2):
                flush.fill_(i & 0xFF)
                torch.cuda.synchronize()
                if i >= 2:
                    kev[i - 2][0].record()
                sp.grad_kernels(buf_parity, run.scale, ws_, st)
                if i >= 2:
                    kev[i - 2][1].record()
            torch.cuda.synchronize()
            kms = sum(e0.elapsed_time(e1) for e0, e1 in kev) / reps
        loss.backward()
        fn.eager_backward = True
        if kms is not None:
            flops = 4.0 * b * N * d * sides  # two of the reference's four backward GEMMs per side
            achieved = flops / (kms * 1e-3) / 1e12
            peak = float(peaks["bf16_tflops"])
            traffic, traffic_src = ncu_traffic("sim_grad_bwd_pair_kernel", b, N, d)
            step_tf = 12.0 * b * N * d * P / (ms_per_step * 1e-3) / 1e12
            roof = {"bound": "tensor",
                    "kernel": "sim_grad_bwd_pair_kernel (the step's fused launch: "
                              f"{sides} (pair, side) problems)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "traffic": traffic * sides if traffic else None, "traffic_source": traffic_src,
                    "algorithmic_bytes": ((b + N) * d * 2 + b * d * 4) * sides,
                    "ms_per_launch": kms, "peak_source": f"{peaks['_source']} burst cuBLAS bf16",
                    "timed": "mmcl_step_grad_kernels on the step's own saved state, L2 flushed "
                             "before every launch",
                    "step_tflops": step_tf, "step_frac_of_burst": step_tf / peak,
                    "step_frac_of_sustained": step_tf / float(peaks["bf16_tflops_sustained"]),
                    "executed_mma_flops_per_step": 12.0 * b * N * d * P,
                    "algorithmic_flops_per_step": 12.0 * b * N * d * P}

    if rank == 0 and world > 1:
        # several ranks: one side of the pair through the ABI entry point (same kernel, one
        # problem, accumulate mode), rank 0 alone, on this rank's share of the batch
        hnd = get_handle(dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        ma, mb = pair_cfg[0][0], pair_cfg[0][1]
        xa_all = full_bf16[ma].to(dev)
        y = full_bf16[mb].to(dev)
        x = xa_all[:b].contiguous()
        sc = torch.tensor([scales[0]], device=dev)
        xinv_all = 1.0 / xa_all.float().norm(dim=1).clamp_min(1e-12)
        xinv = xinv_all[:b].contiguous()
        yinv = 1.0 / y.float().norm(dim=1).clamp_min(1e-12)
        lse_row = torch.empty(b, device=dev)
        lse_col = torch.empty(N, device=dev)
        wsb = lib.mmcl_sim_lse_workspace_bytes(hnd, N, N, d)
        wsf = torch.empty(wsb, dtype=torch.uint8, device=dev)
        check(lib.mmcl_sim_lse_fwd(hnd, ptr(x), b, ptr(y), N, d, ptr(sc), ptr(xinv), ptr(yinv),
                                   ptr(lse_row), None, ptr(wsf), wsb, st), "fwd")
        check(lib.mmcl_sim_lse_fwd(hnd, ptr(y), N, ptr(xa_all), N, d, ptr(sc), ptr(yinv),
                                   ptr(xinv_all), ptr(lse_col), None, ptr(wsf), wsb, st), "fwd")
        dx = torch.empty(b, d, device=dev)
        reps = max(5, min(args.steps, 20))
        kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
               for _ in range(reps)]
        for i in range(reps + 2):
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            if i >= 2:
                kev[i - 2][0].record()
            check(lib.mmcl_sim_grad_bwd_lse(hnd, ptr(x), b, ptr(y), N, d, ptr(sc), ptr(xinv),
                                            ptr(yinv), ptr(lse_row), 2.0 * b, ptr(lse_col), 2.0 * b,
                                            ptr(dx), 1, st), "bwd")
            if i >= 2:
                kev[i - 2][1].record()
        torch.cuda.synchronize()
        kms = sum(e0.elapsed_time(e1) for e0, e1 in kev) / reps
        achieved = 4.0 * b * N * d / (kms * 1e-3) / 1e12
        peak = float(peaks["bf16_tflops"])
        step_tf = 12.0 * b * N * d * P / (ms_per_step * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "sim_grad_bwd_pair_kernel (one side, ABI call)",
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "traffic": None, "algorithmic_bytes": (b + N) * d * 2 + b * d * 4,
                "ms_per_launch": kms, "peak_source": f"{peaks['_source']} burst cuBLAS bf16",
                "timed": "mmcl_sim_grad_bwd_lse (memset + kernel) on rank 0's share with this "
                         "problem's own log-sum-exp vectors, L2 flushed before every launch",
                "step_tflops": step_tf, "step_frac_of_burst": step_tf / peak,
                "step_frac_of_sustained": step_tf / float(peaks["bf16_tflops_sustained"]),
                "executed_mma_flops_per_step": 12.0 * b * N * d * P,
                "algorithmic_flops_per_step": 12.0 * b * N * d * P}
        del xa_all, y, x, dx, wsf

    # ---- (5) CPU baseline beside it (rank 0, N=1 only): bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle.cpu_reference import time_rank_share

        r = time_rank_share(N, d, args.cpu_rows, steps=3, warmup=1, threads=os.cpu_count())
        cpu = {"value": r["samples_per_s"] / P, "unit": "samples/s", "cores": r["threads"],
               "kind": "port",
               "sample": f"one rank's share: {args.cpu_rows} query rows x all {N} keys, both "
                         f"directions, fwd+bwd, 3 steps, torch CPU fp32 ({r['ms_per_step']:.0f} ms/step)"}
        # SURVEY.md section 8d "CPU baseline beside it": BASELINE config 1 whole (N=256, d=512,
        # fp32, W=1) through the oracle's restatement of ContrastiveLoss.forward incl. id matching
        try:
            from oracle.contrastive_oracle import LossPairSpec as OPair
            from oracle.contrastive_oracle import contrastive_oracle, synthetic_inputs

            e1, i1 = synthetic_inputs(256, 512)
            for _ in range(3):
                contrastive_oracle(e1, i1, 1 / 0.07, [OPair(("rgb", "text"))], l2_normalize=True,
                                   local_loss=True, gather_with_grad=True)
            t0 = time.perf_counter()
            for _ in range(30):
                contrastive_oracle(e1, i1, 1 / 0.07, [OPair(("rgb", "text"))], l2_normalize=True,
                                   local_loss=True, gather_with_grad=True)
            dt = (time.perf_counter() - t0) / 30
            cpu["config1"] = {"workload": "BASELINE config 1: N=256, d=512, fp32, W=1, whole "
                                          "forward+backward incl. id matching (oracle port)",
                              "ms_per_step": dt * 1e3, "samples_per_s": 256 / dt}
        except Exception as exc:  # noqa: BLE001
            cpu["config1"] = {"error": str(exc)[:200]}

    # ---- (6) the library baseline on the same GPU: the reference's torch ops (cuBLAS + ATen) -----
    lib_gpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and N <= 32768:
        from oracle.cpu_reference import time_rank_share

        del flush
        torch.cuda.empty_cache()
        try:
            r = time_rank_share(N, d, N, steps=5, warmup=2, device=str(dev), autocast_bf16=True)
            lib_gpu = {"value": r["samples_per_s"] / P, "unit": "samples/s",
                       "ms_per_step": r["ms_per_step"] * P,
                       "kind": "port of the reference's torch ops on cuda:0 under bf16 autocast "
                               "(cuBLAS GEMMs + ATen cross_entropy; logits materialised in HBM); "
                               "id matching excluded (the reference's needs a 2.1 GB N x N temporary)"}
        except Exception as exc:  # noqa: BLE001 - a baseline must never break the bench line
            lib_gpu = {"value": None, "error": str(exc)[:200]}

    if rank == 0:
        line = {
            "metric": metric_name(args.config, N, d),
            "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak" if weak else "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(args.config, N, d, P, world, weak),
                       "rows_per_gpu": b, "l2": "flushed between timed iterations (256 MiB fill kernel on the stream, "
                             "outside each iteration's event pair; no host sync between iterations)",
                       "ms_per_step_incl_flush": ms_per_step_with_flush,
                       "gc": "Python's cyclic garbage collector is run before and disabled inside "
                             "the timed regions",
                       "parallelism": f"dp{world}"},
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "host_ms_per_step": e2e_host,
                    "with_deferred_verification": {
                        "value": e2e_deferred, "host_ms_per_step": e2e_deferred_host,
                        "note": "same loop with ContrastiveLoss.deferred_verification = True "
                                "(opt-in: no host wait inside forward; checked at the end of "
                                "backward)"},
                    "how": "one event pair over all K steps incl. H2D copies, L2 flushes, kernels "
                           "and the D2H copy of loss + d(log scale) every step; the next step's "
                           "H2D is prefetched on a copy stream (double-buffered inputs); the host "
                           "reads each step's loss one step late (after enqueueing the next step) "
                           "and the last one inside the timed region"},
            "gpu_launches": int(launches),
            "tflops_per_gpu": 12.0 * b * N * d * P / (ms_per_step * 1e-3) / 1e12,
            "clocks": clocks,
            "parity": parity,
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
    if parity is not None and not parity["ok"]:
        sys.exit(3)


if __name__ == "__main__":
    main()
