"""Shared-memory layout rules of the kernels, checked on the host (no GPU): tests/data/layout_check.cu includes
rs_device.cuh and verifies alignment, that no two regions that are live together overlap, the two documented aliases
(metric table in the sort's slot arrays, metric denominators in the sort's counters), that the compile-time-shape layouts
equal what rs_create computes for the headline cell (otherwise a handle silently falls back to the general kernel), and
that nine / ten headline cells fit an SM."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="needs nvcc (host-only program, no GPU)")
def test_layout_rules(tmp_path):
    exe = str(tmp_path / "layout_check")
    subprocess.run(["nvcc", "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Wno-deprecated-gpu-targets",
                    "-I", os.path.join(ROOT, "radiosaber_b200", "csrc"), "-o", exe,
                    os.path.join(ROOT, "tests", "data", "layout_check.cu")], check=True, capture_output=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and "layout ok" in r.stdout, r.stdout[-2000:]
