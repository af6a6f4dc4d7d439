#!/usr/bin/env python3
"""Device time of the std::sort emulation alone (development aid): n = 64 x 20 entries whose keys are
the per-(RBG, slice) winner CQIs of the headline workload (max of 5 draws from the trace histogram),
at 1 and at 8 cells per SM, checked against the CPU oracle's real std::sort."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from radiosaber_b200 import sched, workload  # noqa: E402


def main():
    rng = np.random.default_rng(1)
    n = 1280
    for cells in (148, 148 * 8, 148 * 32):
        keys = workload.histogram_cqi(rng, (cells, n, 5)).max(axis=2)
        perm, ms = sched.test_sort_timed(keys, reps=10)
        cyc = ms * 1e-3 * 1.965e9
        waves = -(-cells // (148 * 8))
        print(f"cells {cells:5d}: {ms * 1e3:8.1f} us per launch = {cyc:9.0f} cycles; "
              f"{cells / (ms * 1e-3) / 1e6:7.2f} M sorts/s; per wave {cyc / waves:9.0f} cycles")
    if "--check" in sys.argv:
        from oracle.pyoracle import std_sort_desc
        for i in range(0, cells, 97):
            assert np.array_equal(perm[i], std_sort_desc(keys[i].astype(np.float64))), i
        print("checked against std::sort")


if __name__ == "__main__":
    main()
