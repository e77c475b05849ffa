"""B200-native drop-in for ``mmlearn.modules.losses.ContrastiveLoss``.

Same constructor flags and ``forward(embeddings, example_ids, logit_scale, modality_loss_pairs)``
signature as the reference (mmlearn/modules/losses/contrastive.py:40-65); behind it, PyTorch only
allocates buffers and provides the stream and ``torch.distributed`` plumbing, and every piece of
arithmetic runs in ``libmmcl.so`` (hand-written sm_100a kernels, ``include/mmcl.h``).

Formulation.  For a pair (a, b) with matched, normalised rows A, B and
S = s * A_all B_all^T, rank r owns row block [off_r, off_r + b_r) of S.  Direction-a logits of the
rank are that row block; direction-b logits are the matching *column* block (S is shared by both
directions because the matched global order is the concatenation of the local orders -- checked
at run time).  Forward: two row-wise fused GEMM+LSE passes (A_r x B_all, B_r x A_all).  Backward:
the same two products are recomputed tile by tile with
    G[i,j] = softmax_row(S)[i,j]/den_row + softmax_col(S)[i,j]/den_col(j)
folded into the epilogue, which yields dA_r and dB_r complete on the owning rank: only the [N]
log-sum-exp vectors cross NVLink in backward, never an [N,d] gradient.
"""

from __future__ import annotations

import ctypes
import math
import os
import weakref
from typing import Optional

import torch
import torch.distributed as dist
from torch import nn

from . import _lib
from ._lib import MmclError, check, counts_array, ptr
from .alignment import modality_alignment_loss
from .modalities import LossPairSpec, Modalities
from .plan import PairPlan, all_gather_ints, dist_info
from .steady import SteadyPlan, table_rows

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
_HANDLES: dict = {}
_GENERAL = object()  # _forward_steady: "take the general path" (decided identically on all ranks)


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def get_handle(device: torch.device):
    """One mmcl handle per (device, world).  For world > 1 rank 0 creates the NCCL unique id and
    it is broadcast through the caller's torch.distributed group."""
    import ctypes

    world, rank = dist_info()
    key = (device.index if device.index is not None else torch.cuda.current_device(), world)
    h = _HANDLES.get(key)
    if h is not None:
        return h
    lib = _lib.load()
    uid = None
    if world > 1:
        buf = ctypes.create_string_buffer(_lib.NCCL_UID_BYTES)
        if rank == 0:
            check(lib.mmcl_nccl_unique_id(buf), "mmcl_nccl_unique_id")
        obj = [bytes(buf.raw)]
        dist.broadcast_object_list(obj, src=0)
        uid = ctypes.create_string_buffer(obj[0], _lib.NCCL_UID_BYTES)
    out = ctypes.c_void_p()
    check(lib.mmcl_create(ctypes.byref(out), key[0], rank, world, uid), "mmcl_create")
    _HANDLES[key] = out
    return out


def _validate_ids(first, second) -> None:
    """Error behaviour of mmlearn/datasets/core/example.py:140-157."""
    if not isinstance(first, torch.Tensor) or not isinstance(second, torch.Tensor):
        raise TypeError(
            f"Expected inputs to be tensors, but got {type(first)} and {type(second)}."
        )
    if not (first.ndim == 2 and first.shape[1] == 2):
        raise ValueError(
            "Expected argument `first_example_ids` to be a tensor of shape (N, 2), "
            f"but got shape {first.shape}."
        )
    if not (second.ndim == 2 and second.shape[1] == 2):
        raise ValueError(
            "Expected argument `second_example_ids` to be a tensor of shape (N, 2), "
            f"but got shape {second.shape}."
        )


class _Match:
    """An in-flight device id match; ``info`` = [k, identity, n, m] (device int64).

    ``quick=True`` only tests whether the match is the identity (``mmcl_identity_check``):
    ``info`` is then [n, 1, n, m] or [0, 0, n, m] and no index lists exist."""

    def __init__(self, h, ids_a: torch.Tensor, ids_b: torch.Tensor, capacity: Optional[int] = None,
                 info: Optional[torch.Tensor] = None, quick: bool = False):
        lib = _lib.load()
        self.h, self.ids_a, self.ids_b, self.quick = h, ids_a, ids_b, quick
        dev = ids_a.device
        n, m = ids_a.shape[0], ids_b.shape[0]
        self.info = info if info is not None else torch.empty(4, dtype=torch.long, device=dev)
        if quick:
            self.capacity = 0
            ws_bytes = lib.mmcl_identity_check_workspace_bytes(m)
            self.ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
            check(lib.mmcl_identity_check(h, ptr(ids_a), n, ptr(ids_b), m, ptr(self.info),
                                          ptr(self.ws), self.ws.numel(), _stream(dev)),
                  "mmcl_identity_check")
            return
        self.capacity = max(n, m, 1) if capacity is None else capacity
        self.out_a = torch.empty(self.capacity, dtype=torch.long, device=dev)
        self.out_b = torch.empty(self.capacity, dtype=torch.long, device=dev)
        ws_bytes = lib.mmcl_match_workspace_bytes(n, m)
        self.ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        check(lib.mmcl_match_ids(h, ptr(ids_a), n, ptr(ids_b), m, ptr(self.out_a), ptr(self.out_b),
                                 self.capacity, ptr(self.info), ptr(self.ws), self.ws.numel(),
                                 _stream(dev)), "mmcl_match_ids")

    def finish(self, k: int):
        """Index tensors once the host knows k (re-runs with a larger buffer on overflow)."""
        if k > self.capacity:
            again = _Match(self.h, self.ids_a, self.ids_b, capacity=k)
            return again.out_a[:k], again.out_b[:k]
        return self.out_a[:k], self.out_b[:k]


def find_matching_indices(first_example_ids: torch.Tensor, second_example_ids: torch.Tensor):
    """Device version of ``mmlearn.datasets.core.find_matching_indices``
    (mmlearn/datasets/core/example.py:101-166): same order, same errors."""
    _validate_ids(first_example_ids, second_example_ids)
    if not first_example_ids.is_cuda:
        raise ValueError("mmlearn_b200.find_matching_indices needs CUDA tensors (no CPU fallback)")
    a = first_example_ids.to(torch.long).contiguous()
    b = second_example_ids.to(device=a.device, dtype=torch.long).contiguous()
    m = _Match(get_handle(a.device), a, b)
    k = int(m.info[0].item())
    return m.finish(k)


class _Raw:
    """Raw device memory as a __cuda_array_interface__ object (uint8)."""

    def __init__(self, ptr_: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1",
                                         "data": (ptr_, False), "version": 2}


class _ArenaState:
    """Host-side bookkeeping of a handle's symmetric arena (include/mmcl.h, mmcl_symm_*): which
    buffer (parity) the next forward uses and which run still reads each buffer."""

    def __init__(self):
        self.step = 0
        self.capacity = 0          # bytes per buffer of the current allocation
        self.owner = [None, None]  # weakref to the _Run whose saved tensors live in the buffer
        self.disabled = os.environ.get("MMCL_NO_SYMM", "0") == "1"
        self.views: dict = {}


_ARENAS: dict = {}


class _Transport:
    """How one forward moves data between ranks: direct peer stores into the symmetric arena
    (NVLink, one fence kernel, no NCCL launch) when it is available, NCCL all-gathers otherwise.
    Every decision in here depends only on state that is identical on all ranks."""

    def __init__(self, run, need_bytes: int):
        self.run, self.lib, self.h, self.dev = run, _lib.load(), run.h, run.dev
        self.world, self.rank = dist_info()
        self.st = _stream(self.dev)
        self.arena = False
        self.pending = False
        if self.world == 1:
            return
        key = self.h.value
        a = _ARENAS.setdefault(key, _ArenaState())
        if a.disabled:
            return
        parity = a.step & 1

        def busy(ref):
            """A buffer is claimed by a run that wants gradients and has not been back-propagated
            (or released) yet.  All of that is protocol state every rank shares as long as the
            ranks call forward / backward symmetrically, which DDP training does."""
            prev = ref() if ref is not None else None
            return (prev is not None and prev is not run and bool(prev.states)
                    and prev.backward_calls == 0)

        if busy(a.owner[parity]):
            return  # that run still reads this buffer: use NCCL for this step
        need = max(need_bytes, 1 << 20)
        if a.capacity and need > a.capacity and busy(a.owner[parity ^ 1]):
            # growing re-allocates BOTH buffers; the other one still holds a run's saved state
            return
        if self.lib.mmcl_symm_reserve(self.h, need, self.st) != 0:
            a.disabled = True  # agreed on by all ranks inside the call
            return
        out, nbytes = ctypes.c_void_p(), ctypes.c_size_t()
        check(self.lib.mmcl_symm_buffer(self.h, parity, ctypes.byref(out), ctypes.byref(nbytes)),
              "mmcl_symm_buffer")
        vkey = (parity, out.value, nbytes.value)
        if vkey not in a.views:
            a.views = {k_: v for k_, v in a.views.items() if k_[0] != parity}
            a.views[vkey] = torch.as_tensor(_Raw(out.value, nbytes.value), device=self.dev)
        self.base = a.views[vkey]
        self.capacity = a.capacity = nbytes.value
        self.parity = parity
        self.cursor = 0
        self.arena = True
        a.step += 1
        a.owner[parity] = weakref.ref(run)
        run.arena_claim = (a, parity)

    # ---- arena primitives ----
    def alloc(self, nbytes: int) -> int:
        off = self.cursor
        self.cursor = (off + nbytes + 1023) & ~1023
        if self.cursor > self.capacity:
            raise MmclError("symmetric arena overflow (internal sizing error)")
        return off

    def view(self, off: int, count: int, dtype, shape=None):
        t = self.base[off:off + count * dtype.itemsize].view(dtype)
        return t.view(shape) if shape is not None else t

    def put(self, src: torch.Tensor, nbytes: int, dst_off: int) -> None:
        if nbytes > 0:
            check(self.lib.mmcl_symm_put(self.h, self.parity, ptr(src), nbytes, dst_off, self.st),
                  "mmcl_symm_put")
        self.pending = True

    def fence(self) -> None:
        if self.arena and self.pending:
            check(self.lib.mmcl_symm_fence(self.h, self.st), "mmcl_symm_fence")
            self.pending = False

    # ---- gathers used by the loss ----
    def gather_vector(self, loc: Optional[torch.Tensor], counts, itemsize: int, dtype, row_shape=()):
        """All-gather of per-rank row blocks (counts[r] rows of `itemsize` bytes); returns the
        gathered tensor.  With the arena the data is visible after the next fence()."""
        n = sum(counts)
        off_rows = sum(counts[:self.rank])
        nloc = 0 if loc is None else min(loc.shape[0], counts[self.rank])
        per = itemsize // dtype.itemsize
        if self.arena:
            off = self.alloc(n * itemsize)
            self.put(loc, nloc * itemsize, off + off_rows * itemsize)
            return self.view(off, n * per, dtype, (n, *row_shape) if row_shape else (n,))
        full = torch.empty((n, *row_shape) if row_shape else (n,), dtype=dtype, device=self.dev)
        if nloc:
            full[off_rows:off_rows + nloc].copy_(loc[:nloc])
        check(self.lib.mmcl_all_gather_rows(self.h, full.data_ptr() + off_rows * itemsize, ptr(full),
                                            counts_array(counts), itemsize, self.st),
              "mmcl_all_gather_rows")
        return full


class _PairState:
    """Everything one modality pair needs between forward and backward."""

    __slots__ = ("plan", "key_a", "key_b", "idx_a", "idx_b", "fa_all", "fb_all", "inv_a", "inv_b",
                 "lse_a", "lse_b", "lse_a_all", "lse_b_all", "scale_index", "pos", "dacc")


class _FusedContrastiveFn(torch.autograd.Function):
    """loss = f(logit_scale, *embeddings); all other state travels in ``ctx.run``."""

    @staticmethod
    def forward(ctx, run, logit_scale, *embs):
        loss, dscale = run.forward(logit_scale, embs)
        ctx.run = run
        ctx.save_for_backward(dscale)
        return loss

    @staticmethod
    def backward(ctx, gout):
        (dscale,) = ctx.saved_tensors
        run = ctx.run
        grads = run.backward(gout)
        gs = None
        if ctx.needs_input_grad[1]:
            gs = (dscale * gout).to(run.scale_dtype)
            gs = gs.reshape(run.scale_shape) if run.scale_shape != gs.shape else gs
        return (None, gs, *grads)


class _Run:
    """One forward(+backward) evaluation of the loss."""

    def __init__(self, module, keys, example_ids, pairs):
        self.m = module
        self.keys = keys  # embedding dict keys, in the order the tensors are passed
        self.example_ids = example_ids
        self.pairs = pairs
        self.states: list[_PairState] = []
        self.backward_calls = 0
        self.arena_claim = None  # (_ArenaState, parity) when the saved state lives in the arena

    # ------------------------------------------------------------------------------------
    def forward(self, logit_scale: torch.Tensor, embs):
        m = self.m
        world, rank = dist_info()
        dev = embs[0].device
        self.dev, self.h = dev, get_handle(dev)
        self.embs = dict(zip(self.keys, embs))
        self.scale_shape, self.scale_dtype = logit_scale.shape, logit_scale.dtype
        P = len(self.pairs)
        scale = logit_scale.detach().to(device=dev, dtype=torch.float32).reshape(-1).contiguous()
        if scale.numel() not in (1, P):
            raise ValueError(
                f"logit_scale must be a scalar or have one entry per modality pair ({P}), got "
                f"shape {tuple(logit_scale.shape)}")
        self.scale = scale
        # The host needs the per-rank match counts of every pair to size buffers and order the
        # collectives.  Reading them costs a device->host sync at the very start of the step,
        # with an idle GPU.  When the previous step with the same signature produced a table,
        # the step is enqueued against that table and the table is verified with ONE sync after
        # all forward work has been queued; a mismatch (ids changed structure) re-runs the step.
        # The cache key holds only what every rank shares (pairs, world, train/eval), so all
        # ranks always take the same branch and issue the same collectives with the same sizes;
        # a rank whose local batch no longer fits the table still goes through the motions and
        # the verification (which sees every rank's fresh table) fails on all ranks together.
        sig = (world, tuple(tuple(pr.modalities) for pr in self.pairs), bool(self.need_grad))
        self.fast = None
        assumed = m._plan_cache.get(sig) if m.speculate else None
        if assumed is not None:
            res = _GENERAL
            sp = m._steady.get(sig) if m.native_step else None
            # (shared state only) the steady plan was built for these weights / scale layout
            if sp is not None and sp.shared_key == self._shared_key():
                res = self._forward_steady(sp, assumed)
            if res is _GENERAL:
                res = self._forward_impl(assumed)
            if res is not None:
                return res
            self.states = []  # mis-speculation: fall through to the synchronous path
            self.fast = None
            m._plan_cache.pop(sig, None)
            m._steady.pop(sig, None)
        res = self._forward_impl(None)
        m._plan_cache[sig] = self.all_info
        m._steady[sig] = self._build_steady() if m.native_step else None
        return res

    def _shared_key(self):
        """What a steady plan depends on besides the match table; identical on all ranks."""
        m = self.m
        return (tuple(float(pr.weight) for pr in self.pairs), int(self.scale.numel()),
                m.l2_normalize, m.local_loss, m.gather_with_grad)

    def _build_steady(self) -> Optional[SteadyPlan]:
        """After a successful general step: if its match table describes the steady state (every
        pair everywhere, identity matches, equal row counts), prepare the native step executor
        for the following steps (mmlearn_b200/steady.py)."""
        world, rank = dist_info()
        P = len(self.pairs)
        if P == 0:
            return None
        rows = table_rows(self.all_info.tolist(), world, P)
        if rows is None:
            return None
        if world > 1:
            a = _ARENAS.get(self.h.value)
            if a is None or a.disabled:  # agreed on by all ranks (mmcl_symm_reserve)
                return None
        names, pair_mods = [], []
        for pair in self.pairs:
            idx = []
            for name in pair.modalities:
                nm = Modalities.get_modality(name).name
                if nm not in names:
                    names.append(nm)
                idx.append(names.index(nm))
            pair_mods.append(tuple(idx))
        es = [self.embs.get(Modalities.get_modality(n).embedding) for n in names]
        if any(e is None or e.dtype not in _DTYPES for e in es):
            return None  # cannot happen with this table; stay on the general path
        try:
            sp = SteadyPlan(self.h, world, rank, rows, es[0].shape[-1], names,
                            [_DTYPES[e.dtype] for e in es], [bool(e.requires_grad) for e in es],
                            self.pairs, pair_mods, int(self.scale.numel()), self.m.l2_normalize,
                            self.m.local_loss, self.m.gather_with_grad, bool(self.need_grad))
        except MmclError:
            # outside the fused step's limits (more than 6 modalities / 4 pairs, rows % 4 with
            # several ranks): shapes are rank-identical, so every rank stays on the general path
            return None
        sp.shared_key = self._shared_key()
        return sp

    def _forward_steady(self, sp: SteadyPlan, assumed: torch.Tensor):
        """Speculative step through the native executor; None = mis-speculation, _GENERAL = the
        symmetric arena is not available for this step (same decision on every rank)."""
        m = self.m
        world, rank = dist_info()
        dev = self.dev
        tr = _Transport(self, sp.arena_bytes)
        if world > 1 and not tr.arena:
            return _GENERAL
        for pair in self.pairs:  # same argument errors as the general path / the reference
            ma = Modalities.get_modality(pair.modalities[0])
            mb = Modalities.get_modality(pair.modalities[1])
            if ma.name in self.example_ids and mb.name in self.example_ids:
                _validate_ids(self.example_ids[ma.name], self.example_ids[mb.name])
        local_ok, es, ids = sp.local_inputs(self.embs, self.example_ids, dev)
        if local_ok:
            local_ok = [bool(self.embs[md.embedding].requires_grad) for md in sp.mods] == [
                bool(x) for x in sp.needs_grad]
        ws = torch.empty(sp.ws_bytes, dtype=torch.uint8, device=dev)
        parity = tr.parity if world > 1 else 0
        launch_grad = bool(self.need_grad and m.eager_backward)
        self.states = [sp]  # claims the arena buffer until backward() / release()
        m._resolve_deferred()  # (a plan holds one verification at a time)
        deferred = bool(m.deferred_verification)
        if not sp.forward(parity, local_ok, es, ids, self.scale, ws, launch_grad, _stream(dev),
                          verify=not deferred):
            return None
        if deferred:
            m._deferred = (self, sp)
        self.fast = (sp, parity, es, ws)
        m.steady_steps += 1
        self.grad_kernels_done = launch_grad
        self.all_info = assumed
        self.losses_present = True
        P = sp.P
        out2 = ws[sp.off_out2:sp.off_out2 + 8 * P].view(torch.float32).view(P, 2)
        if P == 1:
            return out2[0, 0], (out2[0, 1:2] if self.scale.numel() > 1 else out2[0, 1])
        return out2[:, 0].sum(), (out2[:, 1].clone() if self.scale.numel() > 1
                                  else out2[:, 1].sum())

    def release(self):
        """Drop the saved state (and with it the claim on the symmetric arena buffer)."""
        self.states = []
        self.fast = None

    def _forward_impl(self, assumed: Optional[torch.Tensor]):
        lib = _lib.load()
        m = self.m
        world, rank = dist_info()
        dev, h = self.dev, self.h
        st = _stream(dev)
        P = len(self.pairs)
        scale = self.scale
        d = next(iter(self.embs.values())).shape[-1]

        # ---- 1. local id matching (contrastive.py:268-291), all pairs in flight at once ----
        ids = {}

        def ids_of(name):
            if name not in ids:
                t = self.example_ids[name]
                ids[name] = t.to(device=dev, dtype=torch.long).contiguous()
            return ids[name]

        # the table as plain Python ints: tensor element access costs microseconds apiece
        assumed_l = assumed.tolist() if assumed is not None else None
        matches: list[Optional[_Match]] = []
        mods = []
        # per pair: [k, identity, rows of ids_a, rows of ids_b], written by the matcher itself
        info = torch.zeros(P, 4, dtype=torch.long, device=dev)
        for p, pair in enumerate(self.pairs):
            ma = Modalities.get_modality(pair.modalities[0])
            mb = Modalities.get_modality(pair.modalities[1])
            mods.append((ma, mb))
            if ma.embedding in self.embs and mb.embedding in self.embs:
                _validate_ids(self.example_ids[ma.name], self.example_ids[mb.name])
                # under an assumed table that says "identity here", only that is tested
                quick = (assumed is not None and bool(assumed_l[rank][p][1])
                         and assumed_l[rank][p][0] > 0)
                matches.append(_Match(h, ids_of(ma.name), ids_of(mb.name), info=info[p],
                                      quick=quick))
            else:
                matches.append(None)
        # upper bound of what this forward keeps in the symmetric arena (rows + norms of both
        # modalities, ids for the cross-rank check, log-sum-exp vectors, the table itself)
        def arena_bytes(T):
            nb = 4096 + world * P * 32
            for p in range(P):
                n_p = sum(T[r][p][0] for r in range(world))
                nb += 2 * (n_p * d * 2 + n_p * 4) + 2 * n_p * 4 + 16 * 1024
                nb += sum(T[r][p][2] + T[r][p][3] for r in range(world)) * 16
            return nb

        tr = None
        if world == 1:
            tr = _Transport(self, 0)
        elif assumed_l is not None:
            tr = _Transport(self, arena_bytes(assumed_l))

        # one exchange: the table of every rank (replaces contrastive.py:197-206 and the shape
        # gathers of _gather_all_tensors)
        if tr is not None and tr.arena:
            off_tab = tr.alloc(world * P * 32)
            tr.put(info, P * 32, off_tab + rank * P * 32)
            all_info_dev = tr.view(off_tab, world * P * 4, torch.long, (world, P, 4))
        else:
            all_info_dev = all_gather_ints(info.view(-1), world).view(world, P, 4)
        all_info = all_info_dev.cpu() if assumed is None else assumed  # host sync (first step)
        self.all_info = all_info
        T = all_info.tolist() if assumed_l is None else assumed_l
        if tr is None:  # first step: the table has just been read, size the arena from it
            tr = _Transport(self, arena_bytes(T))
        verify: list = []  # (device scalar, expected int) pairs checked at the end (speculation)
        gcache: dict = {}

        # ---- 2. per pair: pack matched rows straight into the gather buffer, all-gather ----
        out2 = torch.zeros(P, 2, dtype=torch.float32, device=dev)
        need_grad = self.need_grad  # decided by the caller: grad mode is off inside Function.forward
        shared: dict = {}
        losses_present = False
        for p, pair in enumerate(self.pairs):
            ma, mb = mods[p]
            counts = [T[r][p][0] for r in range(world)]
            n_total = sum(counts)
            if n_total == 0:
                continue  # nobody has this pair: every rank skips it (contrastive.py:304-316)
            plan = PairPlan(counts=counts, rank=rank, local_loss=m.local_loss,
                            gather_with_grad=m.gather_with_grad, weight=float(pair.weight))
            stt = _PairState()
            stt.plan, stt.key_a, stt.key_b = plan, ma.embedding, mb.embedding
            stt.scale_index = p if scale.numel() > 1 else 0
            k = plan.b_local
            ident = bool(T[rank][p][1]) and k > 0
            stt.idx_a = stt.idx_b = None
            if k > 0 and not ident:
                if matches[p] is not None and not matches[p].quick and (
                        assumed is None or k <= matches[p].capacity):
                    stt.idx_a, stt.idx_b = matches[p].finish(k)
                # else: the assumed table does not fit this rank's batch; the step is queued
                # anyway (same collectives as every other rank) and fails verification
            all_ident = all(bool(T[r][p][1]) or counts[r] == 0 for r in range(world))
            gids = None
            if world > 1:
                quick_global = assumed is not None and all(
                    bool(T[r][p][1]) and T[r][p][0] == T[r][p][2] == T[r][p][3]
                    for r in range(world))
                gids = self._global_ids(ma, mb, ids_of, [T[r][p][2] for r in range(world)],
                                        [T[r][p][3] for r in range(world)], gcache, quick_global, tr)

            def gathered(key, idx, name):
                """Raw bf16 rows of every rank's matched samples, [N, d], plus their fp32
                inverse norms [N] (None without l2_normalize).  The local rows are packed
                straight into their slot of the gather buffer (of every rank, when the symmetric
                arena is in use; otherwise an in-place NCCL all-gather follows)."""
                ck = (key, "identity") if all_ident else None
                if ck is not None and ck in shared:
                    return shared[ck]
                src = self.embs[key] if key in self.embs else None
                have = k > 0 and src is not None and src.shape[0] > 0
                s_c = src.detach().contiguous() if have else None
                if have and s_c.dtype not in _DTYPES:
                    raise ValueError(f"unsupported embedding dtype {s_c.dtype}")
                if tr.arena:
                    off_rows = tr.alloc(n_total * d * 2)
                    off_inv = tr.alloc(n_total * 4) if m.l2_normalize else -1
                    f_all = tr.view(off_rows, n_total * d, torch.bfloat16, (n_total, d))
                    inv_all = tr.view(off_inv, n_total, torch.float32) if m.l2_normalize else None
                    if have:
                        check(lib.mmcl_pack_rows_symm(
                            h, tr.parity, ptr(s_c), _DTYPES[s_c.dtype], s_c.shape[0], d, ptr(idx), k,
                            int(m.l2_normalize), off_rows + plan.offset * d * 2,
                            off_inv + 4 * plan.offset if off_inv >= 0 else -1, st),
                            "mmcl_pack_rows_symm")
                    tr.pending = True
                else:
                    f_all = torch.empty(n_total, d, dtype=torch.bfloat16, device=dev)
                    inv_all = (torch.empty(n_total, dtype=torch.float32, device=dev)
                               if m.l2_normalize else None)
                    if have:
                        check(lib.mmcl_pack_rows(
                            h, ptr(s_c), _DTYPES[s_c.dtype], s_c.shape[0], d, ptr(idx), k,
                            int(m.l2_normalize), f_all.data_ptr() + plan.offset * d * 2,
                            inv_all.data_ptr() + 4 * plan.offset if inv_all is not None else None,
                            st), "mmcl_pack_rows")
                    if world > 1:
                        check(lib.mmcl_all_gather_rows(h, f_all.data_ptr() + plan.offset * d * 2,
                                                       ptr(f_all), counts_array(counts), d * 2, st),
                              "mmcl_all_gather_rows")
                        if inv_all is not None:
                            check(lib.mmcl_all_gather_rows(h, inv_all.data_ptr() + 4 * plan.offset,
                                                           ptr(inv_all), counts_array(counts), 4, st),
                                  "mmcl_all_gather_rows")
                if ck is not None:
                    shared[ck] = (f_all, inv_all)
                return f_all, inv_all

            stt.fa_all, stt.inv_a = gathered(ma.embedding, stt.idx_a, ma.name)
            stt.fb_all, stt.inv_b = gathered(mb.embedding, stt.idx_b, mb.name)
            tr.fence()  # arena: every rank's rows, norms, ids and table row are visible from here
            if world > 1:
                kg = self._global_match_count(ma, mb, gids, gcache)
                if assumed is None:
                    self._raise_if_cross_rank(ma, mb, int(kg.item()), n_total)
                else:
                    verify.append((kg, n_total))
            self.states.append(stt)
            if m.cache_labels:
                m._remember_labels(plan, dev)
            if k == 0:
                continue  # skipped process: contributes keys only (contrastive.py:283-301)
            losses_present = True

            # ---- 3. fused similarity + log-sum-exp, both directions ----
            off = plan.offset
            fa = stt.fa_all[off:off + k]
            fb = stt.fb_all[off:off + k]
            sp = scale.data_ptr() + 4 * stt.scale_index
            vec = torch.empty(5, k, dtype=torch.float32, device=dev)
            lse_a, lse_b, es_a, es_b, pos = vec[0], vec[1], vec[2], vec[3], vec[4]
            ws_bytes = lib.mmcl_sim_lse_workspace_bytes(h, k, n_total, d)
            ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
            ia_loc = stt.inv_a[off:off + k] if stt.inv_a is not None else None
            ib_loc = stt.inv_b[off:off + k] if stt.inv_b is not None else None
            check(lib.mmcl_sim_lse_fwd(h, ptr(fa), k, ptr(stt.fb_all), n_total, d, sp, ptr(ia_loc),
                                       ptr(stt.inv_b), ptr(lse_a), ptr(es_a), ptr(ws), ws_bytes, st),
                  "mmcl_sim_lse_fwd")
            check(lib.mmcl_sim_lse_fwd(h, ptr(fb), k, ptr(stt.fa_all), n_total, d, sp, ptr(ib_loc),
                                       ptr(stt.inv_a), ptr(lse_b), ptr(es_b), ptr(ws), ws_bytes, st),
                  "mmcl_sim_lse_fwd")
            check(lib.mmcl_positive_logits(h, ptr(fa), k, ptr(stt.fb_all), n_total, d, off, sp,
                                           ptr(ia_loc), ptr(stt.inv_b), ptr(pos), st),
                  "mmcl_positive_logits")
            check(lib.mmcl_pair_loss(h, ptr(lse_a), ptr(lse_b), ptr(es_a), ptr(es_b), ptr(pos), k,
                                     plan.weight, float(plan.loss_rows), sp,
                                     out2.data_ptr() + 8 * p, st), "mmcl_pair_loss")
            stt.lse_a, stt.lse_b, stt.pos = lse_a, lse_b, pos

        # global loss: every rank reports the mean over all N rows (contrastive.py:334-337)
        if world > 1 and not m.local_loss:
            check(lib.mmcl_all_reduce_sum_f32(h, ptr(out2), out2.numel(), st),
                  "mmcl_all_reduce_sum_f32")
            losses_present = any(s_.plan.n_total > 0 for s_ in self.states)

        # backward needs the log-sum-exp of every row of S (column terms).  Gathered here, in
        # forward, so that backward issues no collective at all (a rank whose loss is a constant
        # 0 never runs backward; the reference deadlocks in that situation).
        if need_grad and world > 1:
            if not tr.arena:
                check(lib.mmcl_group_start(h), "mmcl_group_start")
            for stt in self.states:
                plan = stt.plan
                if plan.col_denominators() is None:
                    continue
                for name in ("lse_a", "lse_b"):
                    loc = getattr(stt, name, None) if plan.b_local > 0 else None
                    setattr(stt, name + "_all",
                            tr.gather_vector(loc, list(plan.counts), 4, torch.float32))
            if not tr.arena:
                check(lib.mmcl_group_end(h), "mmcl_group_end")
        tr.fence()  # also covers a table-only step (no pair had rows)
        self.losses_present = losses_present
        if P == 1:
            loss, dscale = out2[0, 0], (out2[0, 1:2] if scale.numel() > 1 else out2[0, 1])
        else:
            loss = out2[:, 0].sum()
            dscale = out2[:, 1].clone() if scale.numel() > 1 else out2[:, 1].sum()

        self.grad_kernels_done = False
        if assumed is not None:
            # ---- verification of the assumed table: the step's only host wait ----
            # The fresh table is copied to pinned memory behind the forward work; the gradient
            # kernels are queued BEHIND that copy, so the device keeps working while the host
            # waits for the copy (not for the whole stream), returns the loss and runs autograd.
            got_dev = torch.cat([all_info_dev.reshape(-1)] + [kg.reshape(1) for kg, _ in verify])
            got = torch.empty(got_dev.shape, dtype=torch.long, pin_memory=True)
            got.copy_(got_dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            if need_grad and m.eager_backward:
                self.launch_grad_kernels()
            ev.synchronize()
            want = torch.cat([assumed.reshape(-1),
                              torch.tensor([e for _, e in verify], dtype=torch.long)])
            if not torch.equal(got, want):
                return None
        return loss, dscale

    # ------------------------------------------------------------------------------------
    def _global_ids(self, ma, mb, ids_of, sizes_a, sizes_b, cache, quick: bool, tr):
        """Gather the example ids the cross-rank check needs (see _global_match_count); with the
        symmetric arena they become visible at the next fence."""
        key = ("gids", ma.name, mb.name)
        if key in cache:
            return cache[key]
        dev = self.dev

        def gather_ids(name, sizes):
            ck = ("ids", name)
            if ck not in cache:
                loc = (ids_of(name) if name in self.example_ids
                       else torch.empty(0, 2, dtype=torch.long, device=dev))
                cache[ck] = tr.gather_vector(loc, [int(x) for x in sizes], 16, torch.long, (2,))
            return cache[ck]

        # quick: every rank's local match is the identity (verified through the table), so the
        # global match is the concatenation iff the gathered ids are all distinct
        gb = gather_ids(mb.name, sizes_b)
        ga = gb if quick else gather_ids(ma.name, sizes_a)
        cache[key] = (ga, gb, quick)
        return cache[key]

    def _global_match_count(self, ma, mb, gids, cache) -> torch.Tensor:
        """Number of matches between the gathered id lists (device scalar, no host sync).
        The reference matches the *gathered* ids again (contrastive.py:310-313); the fused path
        needs the result only to confirm that the global match list is the concatenation of the
        local ones (true unless an example id occurs on two ranks)."""
        key = ("gcheck", ma.name, mb.name)
        if key not in cache:
            ga, gb, quick = gids
            mt = _Match(self.h, ga, gb, quick=True) if quick else _Match(self.h, ga, gb, capacity=1)
            cache[key] = mt.info[0]
            cache[("keep", key)] = mt  # buffers stay alive until the kernels have run
        return cache[key]

    @staticmethod
    def _raise_if_cross_rank(ma, mb, k_global: int, n_total: int) -> None:
        if k_global != n_total:
            raise NotImplementedError(
                f"pair ({ma.name}, {mb.name}): {k_global} matches across ranks but {n_total} "
                "within ranks -- some example id occurs on more than one rank. The reference's "
                "labels are mis-aligned in that case (SURVEY.md Appendix B); the fused path "
                "refuses it instead of silently computing something else.")

    # ------------------------------------------------------------------------------------
    def launch_grad_kernels(self) -> None:
        """The heavy half of backward: for every pair and side, dacc = sum_j G_ij * yinv_j * Y_j
        (fused recompute kernel).  It depends on nothing that backward() receives, so a
        speculative forward launches it right away: the device works on it while the host waits
        for the table verification and walks through autograd."""
        if self.grad_kernels_done:
            return
        self.grad_kernels_done = True
        lib = _lib.load()
        dev, h = self.dev, self.h
        st = _stream(dev)
        for stt in self.states:
            plan = stt.plan
            k, n, off = plan.b_local, plan.n_total, plan.offset
            stt.dacc = [None, None]
            if k == 0:
                continue
            d = stt.fa_all.shape[1]
            sp = self.scale.data_ptr() + 4 * stt.scale_index
            fa = stt.fa_all[off:off + k]
            fb = stt.fb_all[off:off + k]
            cden = plan.col_denominators()
            ia_loc = stt.inv_a[off:off + k] if stt.inv_a is not None else None
            ib_loc = stt.inv_b[off:off + k] if stt.inv_b is not None else None
            # column-term log-sum-exp vectors: all rows of S (gathered) / all columns of S
            col_lse_for_a = (getattr(stt, "lse_b_all", None) if plan.world > 1 else stt.lse_b)
            col_lse_for_b = (getattr(stt, "lse_a_all", None) if plan.world > 1 else stt.lse_a)
            uniform = cden is None or len(set(cden)) == 1
            if not uniform:
                # uneven ranks: the column denominator depends on the owning rank, so the offset
                # vectors are materialised (mmcl_make_offsets) instead of formed in the kernel
                den = torch.repeat_interleave(
                    torch.tensor(cden, dtype=torch.float32, device=dev),
                    torch.tensor(list(plan.counts), device=dev))

                def make_off(lse, num, dvec, dscalar):
                    o = torch.empty(num, dtype=torch.float32, device=dev)
                    check(lib.mmcl_make_offsets(h, ptr(lse), num, ptr(dvec), dscalar, ptr(o), st),
                          "mmcl_make_offsets")
                    return o

                row_a = make_off(stt.lse_a, k, None, plan.row_denominator())
                row_b = make_off(stt.lse_b, k, None, plan.row_denominator())
                col_for_a = make_off(col_lse_for_a, n, den, 1.0)
                col_for_b = make_off(col_lse_for_b, n, den, 1.0)
            for (key, x, y_all, lse_row, lse_col, inv, yinv, side) in (
                (stt.key_a, fa, stt.fb_all, stt.lse_a, col_lse_for_a, ia_loc, stt.inv_b, 0),
                (stt.key_b, fb, stt.fa_all, stt.lse_b, col_lse_for_b, ib_loc, stt.inv_a, 1),
            ):
                if not self.embs[key].requires_grad:
                    continue
                dacc = torch.empty(k, d, dtype=torch.float32, device=dev)
                if uniform:
                    check(lib.mmcl_sim_grad_bwd_lse(
                        h, ptr(x), k, ptr(y_all), n, d, sp, ptr(inv), ptr(yinv), ptr(lse_row),
                        plan.row_denominator(), ptr(lse_col) if cden is not None else None,
                        cden[0] if cden is not None else 1.0, ptr(dacc), 1, st),
                        "mmcl_sim_grad_bwd_lse")
                else:
                    dacc.zero_()
                    o_row, o_col = (row_a, col_for_a) if side == 0 else (row_b, col_for_b)
                    check(lib.mmcl_sim_grad_bwd(h, ptr(x), k, ptr(y_all), n, d, sp, ptr(inv),
                                                ptr(yinv), ptr(o_row), ptr(o_col), ptr(dacc), st),
                          "mmcl_sim_grad_bwd")
                stt.dacc[side] = dacc

    def _check_saved_state(self) -> None:
        """Second and later backward passes through the same loss (retain_graph=True): the saved
        state is kept until the autograd node dies, but state that lives in the symmetric arena
        may have been overwritten by a later forward, which must not pass silently."""
        self.backward_calls += 1
        if self.backward_calls == 1:
            return
        if not self.states:
            raise RuntimeError(
                "Trying to backward through the fused contrastive loss a second time, but its "
                "saved state has already been released.")
        if self.arena_claim is not None:
            a, parity = self.arena_claim
            owner = a.owner[parity]() if a.owner[parity] is not None else None
            if owner is not self:
                raise RuntimeError(
                    "Trying to backward through the fused contrastive loss a second time after a "
                    "later forward has reused its gather buffer; run the second backward before "
                    "the next-but-one forward of this process.")

    def _backward_steady(self, gout: torch.Tensor):
        self._check_saved_state()
        sp, parity, es, ws = self.fast
        dev = self.dev
        st = _stream(dev)
        if not self.grad_kernels_done:
            sp.grad_kernels(parity, self.scale, ws, st)
            self.grad_kernels_done = True
        gout = gout.detach().to(device=dev, dtype=torch.float32).contiguous()
        by_key = dict(zip((md.embedding for md in sp.mods),
                          sp.backward(parity, es, self.scale, ws, gout, st)))
        if self.m._deferred is not None and self.m._deferred[0] is self:
            # deferred verification: checked once the whole backward pass has been enqueued, i.e.
            # before loss.backward() returns and long before an optimizer could use the gradients
            torch.autograd.Variable._execution_engine.queue_callback(self.m._resolve_deferred)
        out = []
        for key in self.keys:
            src = self.embs[key]
            if not src.requires_grad:
                out.append(None)
            elif by_key.get(key) is not None:
                out.append(by_key[key])
            else:
                out.append(torch.zeros_like(src))
        return out

    def backward(self, gout: torch.Tensor):
        if self.fast is not None:
            return self._backward_steady(gout)
        self._check_saved_state()
        lib = _lib.load()
        dev, h = self.dev, self.h
        st = _stream(dev)
        m = self.m
        self.launch_grad_kernels()
        gout = gout.detach().to(device=dev, dtype=torch.float32).contiguous()
        grads32: dict = {}
        direct: dict = {}
        # how many (pair, side) terms feed each embedding tensor: a tensor fed by exactly one
        # term through the identity row map gets its gradient written once, in its own dtype
        n_terms: dict = {}
        for stt in self.states:
            if stt.plan.b_local > 0:
                for key in (stt.key_a, stt.key_b):
                    n_terms[key] = n_terms.get(key, 0) + 1
        for stt in self.states:
            plan = stt.plan
            k, off = plan.b_local, plan.offset
            if k == 0:
                continue
            d = stt.fa_all.shape[1]
            sp = self.scale.data_ptr() + 4 * stt.scale_index
            fa = stt.fa_all[off:off + k]
            fb = stt.fb_all[off:off + k]
            pc = plan.positive_coefficient()
            ia_loc = stt.inv_a[off:off + k] if stt.inv_a is not None else None
            ib_loc = stt.inv_b[off:off + k] if stt.inv_b is not None else None
            for (key, partner, inv, pinv, idx, side) in (
                (stt.key_a, fb, ia_loc, ib_loc, stt.idx_a, 0),
                (stt.key_b, fa, ib_loc, ia_loc, stt.idx_b, 1),
            ):
                src = self.embs[key]
                dacc = stt.dacc[side]
                if not src.requires_grad or dacc is None:
                    continue
                s_c = src.detach().contiguous()
                is_direct = (idx is None and n_terms.get(key, 0) == 1 and k == src.shape[0]
                             and key not in grads32)
                if is_direct:
                    out_t = torch.empty(src.shape, dtype=src.dtype, device=dev)
                    direct[key] = out_t
                    gbuf, gdt = out_t, _DTYPES[src.dtype]
                else:
                    if key not in grads32:
                        grads32[key] = torch.zeros(src.shape, dtype=torch.float32, device=dev)
                    gbuf, gdt = grads32[key], 0
                check(lib.mmcl_unpack_grad(h, ptr(dacc), ptr(partner), ptr(pinv), k, d, plan.weight,
                                           pc, ptr(gout), sp, ptr(s_c), _DTYPES[s_c.dtype], ptr(inv),
                                           int(m.l2_normalize), ptr(idx), ptr(gbuf), gdt,
                                           int(is_direct), st),
                      "mmcl_unpack_grad")
        out = []
        for key in self.keys:
            src = self.embs[key]
            if not src.requires_grad:
                out.append(None)
            elif key in direct:
                out.append(direct[key])
            elif key in grads32:
                out.append(grads32[key].to(src.dtype))
            else:
                out.append(torch.zeros_like(src))
        return out


class ContrastiveLoss(nn.Module):
    """Contrastive Loss (fused B200 implementation).

    Parameters mirror mmlearn/modules/losses/contrastive.py:23-37.

    Parameters
    ----------
    l2_normalize : bool, optional, default=False
        Whether to L2 normalize the features.
    local_loss : bool, optional, default=False
        Whether to calculate the loss locally i.e. ``local_features@global_features``.
    gather_with_grad : bool, optional, default=False
        Whether to gather tensors with gradients.
    modality_alignment : bool, optional, default=False
        Whether to include the modality alignment loss (dense torch branch, see alignment.py).
    cache_labels : bool, optional, default=False
        Whether to cache the labels.
    """

    def __init__(
        self,
        l2_normalize: bool = False,
        local_loss: bool = False,
        gather_with_grad: bool = False,
        modality_alignment: bool = False,
        cache_labels: bool = False,
    ):
        super().__init__()
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.l2_normalize = l2_normalize
        self.modality_alignment = modality_alignment
        # cache state (same attributes as contrastive.py:55-57; not part of state_dict)
        self._prev_num_logits = 0
        self._labels: dict = {}
        # match-count table of the previous step per call signature (see _Run.forward); plain
        # attributes, not buffers: nothing is added to state_dict
        self._plan_cache: dict = {}
        self.speculate = True
        # steady-state steps go through the native step executor (mmlearn_b200/steady.py)
        self._steady: dict = {}
        self.steady_steps = 0  # diagnostic: forwards that went through the native executor
        self.native_step = os.environ.get("MMCL_NO_NATIVE_STEP", "0") != "1"
        # under speculation, queue the gradient kernels at the end of forward (they need nothing
        # from backward()); set to False if the loss is often evaluated without back-propagation
        self.eager_backward = True
        # Opt-in: do not wait for the device-side verification of a native step inside forward().
        # forward() then returns without any host wait (the host can queue the encoders' backward
        # straight away); the verification is checked at the end of loss.backward() -- or at the
        # next forward() for a loss that is never back-propagated -- and a step whose batch did
        # NOT have the assumed structure raises a RuntimeError there instead of being re-run
        # transparently.  For training loops whose batch structure is fixed (drop_last, paired
        # data); the default re-runs such a step through the general path.
        self.deferred_verification = os.environ.get("MMCL_DEFER_VERIFY", "0") == "1"
        self._deferred = None

    def _resolve_deferred(self) -> None:
        """Check the pending deferred verification, if any (see ``deferred_verification``)."""
        pending, self._deferred = self._deferred, None
        if pending is None:
            return
        run, sp = pending
        if not sp.verify():
            self._plan_cache.clear()
            self._steady.clear()
            raise RuntimeError(
                "mmlearn_b200.ContrastiveLoss(deferred_verification=True): the previous step's "
                "batch did not have the structure of the step before it (example ids not "
                "identical row by row in every pair on every rank, or duplicated); its loss and "
                "gradients are invalid. Re-run the step; it will take the general path.")

    def _remember_labels(self, plan: PairPlan, device) -> None:
        """contrastive.py:190-219: rebuild when the logit width changes or the device is new."""
        n = plan.n_total if plan.n_logits is None else plan.n_logits
        if self._prev_num_logits != n or device not in self._labels:
            self._labels[device] = plan.labels(device)
            self._prev_num_logits = n

    def get_ground_truth(self, plan: PairPlan, device) -> torch.Tensor:
        """Labels of one pair on this rank (``_get_ground_truth``, contrastive.py:162-219)."""
        return plan.labels(device)

    def forward(
        self,
        embeddings: dict[str, torch.Tensor],
        example_ids: dict[str, torch.Tensor],
        logit_scale: torch.Tensor,
        modality_loss_pairs: list[LossPairSpec],
    ) -> torch.Tensor:
        """Calculate the contrastive loss (same contract as contrastive.py:59-85)."""
        self._resolve_deferred()
        if not embeddings:
            raise ValueError("embeddings is empty")
        keys = list(embeddings.keys())
        embs = [embeddings[k] for k in keys]
        for k, e in zip(keys, embs):
            if not isinstance(e, torch.Tensor) or not e.is_cuda:
                raise ValueError(
                    f"embedding '{k}' is not a CUDA tensor: mmlearn_b200 runs sm_100a kernels only "
                    "and has no CPU fallback")
            if e.ndim != 2 or e.shape[-1] % 8 != 0 or e.shape[-1] > 2048:
                raise ValueError(
                    f"embedding '{k}' must be [batch, d] with d a multiple of 8 (<= 2048), got "
                    f"{tuple(e.shape)}")
            if e.dtype not in _DTYPES:
                raise ValueError(f"embedding '{k}' has unsupported dtype {e.dtype}")
        if len({e.shape[-1] for e in embs}) != 1:
            raise ValueError("Expected all embeddings to have the same dimension.")
        if not isinstance(logit_scale, torch.Tensor):
            logit_scale = torch.tensor(float(logit_scale), device=embs[0].device)
        run = _Run(self, keys, example_ids, list(modality_loss_pairs))
        run.need_grad = torch.is_grad_enabled() and (
            any(e.requires_grad for e in embs) or logit_scale.requires_grad)
        loss = _FusedContrastiveFn.apply(run, logit_scale, *embs)
        if not run.need_grad:
            run.release()  # nothing will be back-propagated: free the arena buffer for reuse
        if self.modality_alignment:
            # contrastive.py:146-149; a dense torch branch outside the fused path (alignment.py)
            align = modality_alignment_loss(embeddings, logit_scale, self.l2_normalize,
                                            self.local_loss, self.gather_with_grad)
            return loss + align if run.losses_present else align
        if not run.losses_present:
            # no loss to compute (e.g. no paired data in batch): contrastive.py:151-158
            return torch.zeros((), device=logit_scale.device, dtype=embs[0].dtype)
        return loss


FusedContrastiveLoss = ContrastiveLoss

try:  # pragma: no cover - hydra-zen / mmlearn are not installable in the build image
    from mmlearn.conf import external_store  # type: ignore

    external_store(ContrastiveLoss, name="FusedContrastiveLoss", group="modules/losses",
                   provider="mmlearn_b200")
except Exception:  # noqa: BLE001
    pass
