"""Registers of the built kernels (cuobjdump --dump-resource-usage on librs_sched.so; no GPU): the cells an SM holds are
decided by the register count of the instantiation (65536 / 128 threads / cells), so the caps of rs_device.cuh
min_cells_per_sm() and "no spills in the headline kernel" are checked where a regression would otherwise only show as a
slower bench (an edit to the sort once took the headline from nine cells per SM to eight that way)."""
import re
import shutil
import subprocess

import pytest

from radiosaber_b200 import sched


def _usage():
    out = subprocess.run(["cuobjdump", "--dump-resource-usage", sched.LIB_PATH], capture_output=True, text=True).stdout
    res = {}
    for m in re.finditer(r"Function (\S+?):\s*\n\s*REG:(\d+) STACK:(\d+)", out):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        res[name] = (int(m.group(2)), int(m.group(3)))
    return res


@pytest.mark.skipif(shutil.which("cuobjdump") is None or shutil.which("c++filt") is None, reason="needs cuobjdump and c++filt")
def test_register_caps_of_the_compile_time_shape_kernels():
    use = _usage()
    fixed = {k: v for k, v in use.items() if "rs_tti_kernel<" in k and "FixedShape<20, 5, 64, 8" in k}
    assert len(fixed) == 7 * 2 * 2, sorted(fixed)            # seven ids x two CQI layouts x streamed / trace-driven CQI
    for name, (reg, stack) in fixed.items():
        algo, trace = re.search(r"rs_tti_kernel<(\d+), (true|false)", name).groups()
        algo, trace = int(algo), trace == "true"
        ten = algo == 8 or (algo in (101, 103) and not trace)
        cap = 48 if ten else (64 if algo == 11 else 56)
        assert reg <= cap, (name, reg, cap)
        if algo == 9 and not trace:
            assert stack == 0, (name, "the headline kernel spills", stack)
    # the general kernels keep eight cells per SM (64 registers) and at most cold spills
    general = {k: v for k, v in use.items() if "rs::rs_tti_kernel<" in k and "DynShape" in k}
    assert len(general) == 8 * 2 * 2
    for name, (reg, stack) in general.items():
        assert reg <= 64 and stack <= 64, (name, reg, stack)
