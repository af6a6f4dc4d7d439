"""The inputs of the reference's own two unit-test programs (unittest/test_tp_algos.cpp: a 5 RBG x 3 slice CQI
matrix with quotas {1,3,1}; unittest/test_effective_sinr.cpp: seven RBs at 20 dB and one at 8 dB).  The reference
records no expected output, so tests/golden/unittest_vectors.json holds what its functions return when compiled
here (oracle/unittest_probe.cpp, tools/make_unittest_vectors.py)."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import OracleScheduler
from radiosaber_b200 import sched
from tests.helpers import GOLDEN

VEC = json.load(open(os.path.join(GOLDEN, "unittest_vectors.json")))


def _cell(cls, algo):
    """The 5 x 3 matrix as a cell: 5 RBs in RBGs of one, three slices of one UE whose CQI on RBG g is the matrix
    entry, weights 0.2 / 0.6 / 0.2 so that the RBG quotas come out as {1, 3, 1}."""
    m = np.array(VEC["cqi_matrix"], dtype=np.uint8)               # [rbg][slice]
    p = np.tile(np.array([0, 0, 1, 1], dtype=np.int32), (3, 1))
    s = cls(algo, [0.2, 0.6, 0.2], p, np.arange(3, dtype=np.int32), 1, n_rbs=5, rbg_size=1)
    out = s.step(np.ascontiguousarray(m.T)[None], np.zeros((1, 2), np.int32), dt=0.001, want_aux=True)
    assert out["slice_quota"][0].tolist() == VEC["quota"]
    return out["rbg_to_ue"][0].tolist()   # one UE per slice: the UE id is the slice id


def test_oracle_maximize_cell_on_the_reference_unittest_matrix():
    assert _cell(OracleScheduler, 9) == VEC["maximize_cell"]


def test_oracle_eesm_on_the_reference_unittest_vector():
    L = pyoracle.lib()
    sinr = np.array(VEC["sinr_db"], dtype=np.float64)
    eff = L.rso_eesm_effective_sinr(sinr.ctypes.data, len(sinr))
    assert eff == VEC["eesm_effective_sinr"]                      # bit for bit
    cqi = L.rso_cqi_from_sinr(eff)
    assert cqi == VEC["cqi"] and L.rso_mcs_from_cqi(cqi) == VEC["mcs"]
    assert L.rso_tbs_from_mcs(VEC["mcs"], len(sinr), None) == VEC["tbs_8rb"]


@pytest.mark.gpu
def test_cuda_maximize_cell_on_the_reference_unittest_matrix():
    assert _cell(sched.Scheduler, 9) == VEC["maximize_cell"]


@pytest.mark.gpu
@pytest.mark.parametrize("algo", [8, 10, 101, 103])
def test_cuda_other_inter_slice_algorithms_on_the_unittest_matrix(algo):
    """The same tiny cell (5 RBGs, 3 slices: far below the 64 x 20 the scratch buffers are sized for in the other
    tests) under the other inter-slice algorithms: CUDA == oracle."""
    assert _cell(sched.Scheduler, algo) == _cell(OracleScheduler, algo)


@pytest.mark.gpu
def test_cuda_link_adaptation_on_the_reference_unittest_vector():
    """Seven RBs whose CQI maps to >= 20 dB and one at the CQI of 8 dB cannot be fed as dB values (the path takes
    CQI); instead the same vector through CQI: the device EESM/TBS of one UE holding 8 one-RB RBGs equals the oracle's."""
    cqi = np.array([[11, 11, 11, 11, 11, 11, 11, 6]], dtype=np.uint8)
    p = np.array([[0, 0, 1, 1]], dtype=np.int32)
    outs = []
    for cls in (OracleScheduler, sched.Scheduler):
        s = cls(9, [1.0], p, np.zeros(1, np.int32), 1, n_rbs=8, rbg_size=1)
        outs.append(s.step(cqi[None], np.zeros((1, 2), np.int32), dt=0.001, want_aux=True))
    for k in ("rbg_to_ue", "tbs_bits", "mcs", "final_cqi"):
        assert np.array_equal(outs[0][k], outs[1][k]), k
    assert outs[0]["tbs_bits"][0, 0] > 0
