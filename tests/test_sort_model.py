"""CPU: the level-parallel introsort formulation the kernel implements (tests/sort_model.py)
reproduces the real libstdc++ std::sort permutation."""
import numpy as np
import pytest

from oracle import pyoracle
from radiosaber_b200 import workload
from tests.sort_model import level_parallel_introsort


def _cqi_keys(rng, n):
    p = workload.CQI_HIST.astype(np.float64) / workload.CQI_TOTAL
    return rng.choice(np.arange(1, 16), size=n, p=p).astype(np.float64)


@pytest.mark.parametrize("n", [1, 2, 15, 16, 17, 18, 33, 64, 100, 320, 1280, 1281, 3200, 4096])
def test_model_matches_std_sort(n):
    rng = np.random.default_rng(n)
    for trial in range(30 if n <= 1280 else 6):
        kind = trial % 5
        if kind == 0:
            keys = _cqi_keys(rng, n)
        elif kind == 1:
            keys = rng.integers(0, 16, size=n).astype(np.float64)
        elif kind == 2:
            keys = np.full(n, 7.0)
        elif kind == 3:
            keys = np.sort(rng.integers(0, 16, size=n)).astype(np.float64)
        else:
            keys = np.maximum.reduce(_cqi_keys(rng, 5 * n).reshape(5, n))  # per-slice max of 5 UEs
        want = pyoracle.std_sort_desc(keys)
        got = level_parallel_introsort(keys)
        assert np.array_equal(want, got), (n, trial)


@pytest.mark.parametrize("depth", [0, 1, 2, 3, 5])
def test_model_heap_fallback_matches_emulation(depth):
    rng = np.random.default_rng(100 + depth)
    for n in (17, 40, 200, 1280):
        keys = rng.integers(0, 16, size=n).astype(np.float64)
        want = pyoracle.introsort_emul_desc(keys, depth)
        got = level_parallel_introsort(keys, depth)
        assert np.array_equal(want, got), (n, depth)


def test_emulation_matches_std_sort():
    rng = np.random.default_rng(7)
    for n in (5, 64, 1280, 3000):
        keys = rng.integers(0, 16, size=n).astype(np.float64)
        assert np.array_equal(pyoracle.std_sort_desc(keys), pyoracle.introsort_emul_desc(keys, -1))
