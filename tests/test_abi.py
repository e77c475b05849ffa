"""CPU checks of the drop-in boundary: the C-ABI library loads and exports exactly the symbols
include/mmcl.h declares (no compute call is made here)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "mmcl.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mmcl_[a-z0-9_]+)\s*\(", txt)))


def test_header_and_binding_agree():
    from mmlearn_b200 import _lib

    assert _declared() == sorted(_lib.SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    from mmlearn_b200 import _lib

    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__

        __graft_entry__.build()
    lib = _lib.load()
    for name in _declared():
        assert hasattr(lib, name), name
    assert lib.mmcl_version() == 100
    # argument validation happens before any CUDA call: safe without a GPU
    assert lib.mmcl_match_workspace_bytes(10, 10) > 0
    assert lib.mmcl_nccl_unique_id(None) == -1
    assert b"null uid" in lib.mmcl_last_error()
    # the step executor validates before touching CUDA as well
    assert lib.mmcl_step_workspace_bytes(None) == 0 and lib.mmcl_step_arena_bytes(None) == 0
    assert lib.mmcl_step_create(None, None, None) == -1
    assert b"mmcl_step_create" in lib.mmcl_last_error()
    assert lib.mmcl_step_verify(None, None) == -1
    assert lib.mmcl_step_destroy(None) == 0
    assert lib.mmcl_sim_rank_fwd(None, None, 0, None, 0, 8, None, None, None, None, None, None, 0,
                                 None) == -1


def test_no_cpu_fallback():
    import torch

    from mmlearn_b200 import ContrastiveLoss, LossPairSpec

    fn = ContrastiveLoss()
    with pytest.raises(ValueError, match="no CPU fallback"):
        fn({"rgb_embedding": torch.randn(4, 8), "text_embedding": torch.randn(4, 8)},
           {"rgb": torch.zeros(4, 2, dtype=torch.long), "text": torch.zeros(4, 2, dtype=torch.long)},
           torch.tensor(1.0), [LossPairSpec(("rgb", "text"))])


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/mmcl.h must compile as C99 (and as C++) on its own,
    with no torch / CUDA headers."""
    import shutil
    import subprocess

    hdr = os.path.join(ROOT, "include", "mmcl.h")
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++17", "-Wall", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert r.returncode == 0, r.stdout
    txt = open(hdr).read()
    assert "#include <cuda" not in txt and "#include <torch" not in txt
