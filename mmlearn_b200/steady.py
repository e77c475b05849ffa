"""Steady-state fast path of the fused loss: one native call per forward / backward.

When a step has the same structure as the previous one -- every pair present on every rank,
every local id match the identity, the same number of rows on every rank -- the launch sequence of
``loss._Run._forward_impl`` is fixed.  ``SteadyPlan`` hands that sequence to the C step executor
(``mmcl_step_*`` in include/mmcl.h), which enqueues it from native code.  The assumption is
verified on the device every step (fresh id-match table of all ranks + cross-rank duplicate
check); a failed verification makes every rank fall back to the general path together.

Everything that decides *whether* the fast path is taken depends only on state all ranks share
(the assumed table, the call signature, module flags); a rank whose local batch does not fit still
goes through the collective motions (``local_ok = 0``) so that no rank is left waiting.
"""

from __future__ import annotations

import ctypes
import weakref
from ctypes import c_int, c_int32, c_void_p
from typing import Optional

import torch

from . import _lib
from ._lib import StepDesc, StepLayout, StepPair, check
from .modalities import Modalities
from .plan import PairPlan

_DTYPES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def table_rows(table: list, world: int, n_pairs: int) -> Optional[int]:
    """``rows`` if the match table (nested list [world][P][4] = k, identity, n, m) describes the
    steady state (k = n = m = rows > 0 and identity, everywhere), else None."""
    rows = table[0][0][0] if (n_pairs > 0 and table) else 0
    if rows <= 0:
        return None
    want = [rows, 1, rows, rows]
    for r in range(world):
        for p in range(n_pairs):
            if list(table[r][p]) != want:
                return None
    return int(rows)


class SteadyPlan:
    """One ``mmcl_step`` plus the host-side description it was built from."""

    def __init__(self, h, world: int, rank: int, rows: int, d: int, names: list,
                 dtypes: list, needs_grad: list, pair_specs: list, pair_mods: list, n_scales: int,
                 l2_normalize: bool, local_loss: bool, gather_with_grad: bool, need_grad: bool):
        lib = _lib.load()
        self.h, self.world, self.rank, self.rows, self.d = h, world, rank, rows, d
        self.names = names                    # modality names in plan order
        self.mods = [Modalities.get_modality(n) for n in names]
        self.dtypes, self.needs_grad = list(dtypes), list(needs_grad)
        self.P, self.M = len(pair_specs), len(names)
        self.need_grad = need_grad
        counts = [rows] * world
        self.plans = []
        pairs = (StepPair * self.P)()
        for p, (spec, (ia, ib)) in enumerate(zip(pair_specs, pair_mods)):
            plan = PairPlan(counts=counts, rank=rank, local_loss=local_loss,
                            gather_with_grad=gather_with_grad, weight=float(spec.weight))
            self.plans.append(plan)
            cden = plan.col_denominators()
            pairs[p] = StepPair(ia, ib, p if n_scales > 1 else 0, 0 if cden is None else 1,
                                plan.weight, float(plan.loss_rows), plan.row_denominator(),
                                cden[0] if cden is not None else 1.0,
                                plan.positive_coefficient())
        self._pairs = pairs
        self._dt = (c_int32 * self.M)(*self.dtypes)
        self._ng = (c_int32 * self.M)(*[int(x) for x in self.needs_grad])
        desc = StepDesc(self.M, self.P, int(l2_normalize), int(world > 1 and not local_loss),
                        int(need_grad), 0, rows, d, self._dt, self._ng, pairs)
        out = c_void_p()
        check(lib.mmcl_step_create(h, ctypes.byref(desc), ctypes.byref(out)), "mmcl_step_create")
        self.step = out
        self.ws_bytes = lib.mmcl_step_workspace_bytes(out)
        self.arena_bytes = lib.mmcl_step_arena_bytes(out)
        lay = StepLayout()
        check(lib.mmcl_step_layout(out, ctypes.byref(lay)), "mmcl_step_layout")
        self.off_out2, self.off_vec, self.off_info = lay.off_out2, lay.off_vec, lay.off_info
        # how many (pair, side) terms feed each modality: exactly one -> direct stores
        self.n_terms = [0] * self.M
        for ia, ib in pair_mods:
            self.n_terms[ia] += 1
            self.n_terms[ib] += 1
        weakref.finalize(self, lib.mmcl_step_destroy, out)

    # ------------------------------------------------------------------------------------
    def local_inputs(self, embs: dict, example_ids: dict, dev):
        """(local_ok, embedding tensors, id tensors) of this rank for the plan's modalities."""
        es, ids = [], []
        for mod, dt in zip(self.mods, self.dtypes):
            e = embs.get(mod.embedding)
            i = example_ids.get(mod.name)
            if (e is None or not isinstance(i, torch.Tensor) or e.shape != (self.rows, self.d)
                    or _DTYPES.get(e.dtype) != dt or i.ndim != 2 or i.shape != (self.rows, 2)):
                return False, [], []
            es.append(e.detach().contiguous())
            ids.append(i.to(device=dev, dtype=torch.long).contiguous())
        return True, es, ids

    def forward(self, parity: int, local_ok: bool, es: list, ids: list, scale: torch.Tensor,
                ws: torch.Tensor, launch_grad: bool, stream: int, verify: bool = True) -> bool:
        """Enqueue the step and verify the assumption; False = mis-speculation.
        ``verify=False`` returns at once (True); the caller must call :meth:`verify` before it
        trusts the step's results and before the next ``forward`` of this plan."""
        lib = _lib.load()
        if local_ok:
            e_arr = (c_void_p * self.M)(*[t.data_ptr() for t in es])
            i_arr = (c_void_p * self.M)(*[t.data_ptr() for t in ids])
        else:
            e_arr = i_arr = None
        check(lib.mmcl_step_forward(self.step, parity, int(local_ok), e_arr, i_arr,
                                    scale.data_ptr(), ws.data_ptr(), int(launch_grad), stream),
              "mmcl_step_forward")
        return self.verify() if verify else True

    def verify(self) -> bool:
        """Wait for the verification word of the last forward (not for the stream)."""
        ok = c_int(0)
        check(_lib.load().mmcl_step_verify(self.step, ctypes.byref(ok)), "mmcl_step_verify")
        return bool(ok.value)

    def grad_kernels(self, parity: int, scale: torch.Tensor, ws: torch.Tensor, stream: int):
        check(_lib.load().mmcl_step_grad_kernels(self.step, parity, scale.data_ptr(),
                                                 ws.data_ptr(), stream), "mmcl_step_grad_kernels")

    def backward(self, parity: int, es: list, scale: torch.Tensor, ws: torch.Tensor,
                 gout: torch.Tensor, stream: int) -> list:
        """Gradient of every plan modality (None where not required), in the leaf dtype."""
        lib = _lib.load()
        bufs, direct = [], []
        for m in range(self.M):
            if not self.needs_grad[m]:
                bufs.append(None)
                direct.append(0)
            elif self.n_terms[m] == 1:
                bufs.append(torch.empty_like(es[m]))
                direct.append(1)
            else:
                bufs.append(torch.zeros(es[m].shape, dtype=torch.float32, device=es[m].device))
                direct.append(0)
        e_arr = (c_void_p * self.M)(*[t.data_ptr() for t in es])
        g_arr = (c_void_p * self.M)(*[None if t is None else t.data_ptr() for t in bufs])
        d_arr = (c_int32 * self.M)(*direct)
        check(lib.mmcl_step_backward(self.step, parity, e_arr, scale.data_ptr(), ws.data_ptr(),
                                     gout.data_ptr(), g_arr, d_arr, stream), "mmcl_step_backward")
        return [None if t is None else (t if dr else t.to(es[m].dtype))
                for m, (t, dr) in enumerate(zip(bufs, direct))]
