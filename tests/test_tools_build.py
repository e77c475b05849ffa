"""The measurement tools that compile the backward kernel with its timeline instrumentation
(MMCL_TRACE / MMCL_TRACE_FINE are not part of the normal build) must keep compiling for sm_100a."""
import os
import shutil
import subprocess
import tempfile

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17"]


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
@pytest.mark.parametrize("src,flags", [("bwd_trace.cu", ["-DMMCL_TRACE_FINE"]),
                                       ("dsmem_rate.cu", [])])
def test_probe_compiles(src, flags):
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "probe.o")
        r = subprocess.run(["nvcc", *ARCH, *flags, "-c", os.path.join(ROOT, "tools", src), "-o", out],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        assert os.path.getsize(out) > 0
