"""ctypes binding of the C ABI in ``include/mmcl.h`` (``mmlearn_b200/libmmcl.so``).

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""

from __future__ import annotations

import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmmcl.so")



class StepPair(Structure):
    """mmcl_step_pair (include/mmcl.h)."""

    _fields_ = [("mod_a", c_int32), ("mod_b", c_int32), ("scale_index", c_int32),
                ("has_col", c_int32), ("weight", c_double), ("loss_rows", c_double),
                ("row_den", c_double), ("col_den", c_double), ("pos_coef", c_double)]


class StepDesc(Structure):
    """mmcl_step_desc (include/mmcl.h)."""

    _fields_ = [("n_modalities", c_int32), ("n_pairs", c_int32), ("normalize", c_int32),
                ("all_reduce_loss", c_int32), ("need_grad", c_int32), ("reserved", c_int32),
                ("rows", c_int64), ("d", c_int64), ("mod_dtype", POINTER(c_int32)),
                ("mod_needs_grad", POINTER(c_int32)), ("pairs", POINTER(StepPair))]


class StepLayout(Structure):
    """mmcl_step_layout_t (include/mmcl.h)."""

    _fields_ = [("off_out2", c_size_t), ("off_vec", c_size_t), ("off_info", c_size_t)]


#: every symbol include/mmcl.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "mmcl_version": (c_int, []),
    "mmcl_last_error": (c_char_p, []),
    "mmcl_kernel_launches": (c_int64, []),
    "mmcl_nccl_unique_id": (c_int, [c_void_p]),
    "mmcl_create": (c_int, [POINTER(c_void_p), c_int, c_int, c_int, c_void_p]),
    "mmcl_destroy": (c_int, [c_void_p]),
    "mmcl_match_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "mmcl_match_ids": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                               c_int64, c_void_p, c_void_p, c_size_t, c_void_p]),
    "mmcl_identity_check_workspace_bytes": (c_size_t, [c_int64]),
    "mmcl_identity_check": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                    c_void_p, c_size_t, c_void_p]),
    "mmcl_pack_rows": (c_int, [c_void_p, c_void_p, c_int, c_int64, c_int64, c_void_p, c_int64,
                               c_int, c_void_p, c_void_p, c_void_p]),
    "mmcl_symm_reserve": (c_int, [c_void_p, c_size_t, c_void_p]),
    "mmcl_symm_buffer": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_size_t)]),
    "mmcl_symm_put": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_size_t, c_void_p]),
    "mmcl_pack_rows_symm": (c_int, [c_void_p, c_int, c_void_p, c_int, c_int64, c_int64, c_void_p,
                                    c_int64, c_int, c_size_t, c_int64, c_void_p]),
    "mmcl_symm_fence": (c_int, [c_void_p, c_void_p]),
    "mmcl_sim_bce_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int64, c_int64]),
    "mmcl_sim_bce_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "mmcl_sim_bce_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p]),
    "mmcl_sim_logits_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                    c_double, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "mmcl_debug_push": (c_int, [c_void_p, c_int, c_size_t, c_size_t, c_int, c_int, c_int, c_void_p]),
    "mmcl_group_start": (c_int, [c_void_p]),
    "mmcl_group_end": (c_int, [c_void_p]),
    "mmcl_all_gather_rows": (c_int, [c_void_p, c_void_p, c_void_p, POINTER(c_int64), c_int64,
                                     c_void_p]),
    "mmcl_all_reduce_sum_f32": (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "mmcl_sim_lse_workspace_bytes": (c_size_t, [c_void_p, c_int64, c_int64, c_int64]),
    "mmcl_sim_lse_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "mmcl_positive_logits": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                     c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmcl_logit_scale": (c_int, [c_void_p, c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "mmcl_indexed_similarity": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                        c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                        c_void_p]),
    "mmcl_sim_rank_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                  c_void_p]),
    "mmcl_pair_loss": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                               c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "mmcl_sim_grad_bwd": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p]),
    "mmcl_sim_grad_bwd_lse": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_int64, c_int64,
                                      c_void_p, c_void_p, c_void_p, c_void_p, c_double, c_void_p,
                                      c_double, c_void_p, c_int, c_void_p]),
    "mmcl_make_offsets": (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_double, c_void_p,
                                  c_void_p]),
    "mmcl_unpack_grad": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_double,
                                 c_double, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                 c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "mmcl_step_create": (c_int, [c_void_p, POINTER(StepDesc), POINTER(c_void_p)]),
    "mmcl_step_destroy": (c_int, [c_void_p]),
    "mmcl_step_workspace_bytes": (c_size_t, [c_void_p]),
    "mmcl_step_arena_bytes": (c_size_t, [c_void_p]),
    "mmcl_step_layout": (c_int, [c_void_p, POINTER(StepLayout)]),
    "mmcl_step_forward": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p), POINTER(c_void_p),
                                  c_void_p, c_void_p, c_int, c_void_p]),
    "mmcl_step_verify": (c_int, [c_void_p, POINTER(c_int)]),
    "mmcl_step_grad_kernels": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mmcl_step_backward": (c_int, [c_void_p, c_int, POINTER(c_void_p), c_void_p, c_void_p, c_void_p,
                                   POINTER(c_void_p), POINTER(c_int32), c_void_p]),
}

NCCL_UID_BYTES = 128
_lib = None


class MmclError(RuntimeError):
    """A C-ABI call returned a negative mmcl_status."""


def load() -> ctypes.CDLL:
    """Load libmmcl.so and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python __graft_entry__.py` "
            "(nvcc, sm_100a). mmlearn_b200 has no CPU or PyTorch fallback."
        )
    lib = ctypes.CDLL(LIB_PATH, mode=ctypes.RTLD_GLOBAL)
    for name, (restype, argtypes) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().mmcl_last_error()
        raise MmclError(f"{what} failed (status {rc}): {msg.decode() if msg else ''}")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def counts_array(counts) -> ctypes.Array:
    return (c_int64 * len(counts))(*[int(c) for c in counts])
