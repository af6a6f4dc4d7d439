"""GPU: a bounded slice of tools/fuzz_parity.py in the suite -- random slice configurations, shapes, CQI layouts,
idle bearers, finite queues and head-of-line delays, 40-260 TTIs each, CUDA against the CPU oracle (outputs of
every TTI and the cell state at the end, bit-exact).  The soak version is `python tools/fuzz_parity.py`."""
import numpy as np
import pytest

from tools import fuzz_parity

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [101, 202, 303, 404])
def test_random_configurations_match_the_oracle(seed):
    rng = np.random.default_rng(seed)
    for case in range(6):
        err = fuzz_parity.one_case(rng, case)
        assert err is None, f"seed {seed}: {err}"
