"""CPU: the oracle restatement (oracle/rs_oracle.cpp) against the golden vectors recorded from
the unmodified reference (tools/make_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

from oracle.pyoracle import OracleScheduler
from radiosaber_b200 import workload
from tests.helpers import golden_names, load_golden, replay_golden


@pytest.mark.parametrize("name", golden_names())
def test_oracle_matches_reference_record(name):
    rec = load_golden(name)
    o = OracleScheduler(int(rec["algo"]), rec["weight"], rec["params"], rec["ue_to_slice"], 1,
                        n_rbs=int(rec["R"]), rbg_size=int(rec["rbg_size"]), cqi_per_rb=int(rec["cqi_per_rb"]),
                        dead_work=1)
    bad = replay_golden(o, rec)
    assert not bad, bad[:10]


def test_dead_work_flag_changes_nothing():
    rec = load_golden("a9_fix20x5_synth")
    o = OracleScheduler(9, rec["weight"], rec["params"], rec["ue_to_slice"], 1, dead_work=0)
    assert not replay_golden(o, rec)


def test_tti_clock_matches_reference_clock():
    rec = load_golden("a9_fix20x5_trace")
    now, dt = workload.tti_clock(int(rec["T"]))
    assert np.array_equal(now, rec["now"])
    assert np.array_equal(dt, rec["dt"])


def test_golden_covers_out_of_bounds_tbs_row():
    """some UE must hold a multiple of 5 RBGs > 110 RBs so row -1 is exercised."""
    hit = 0
    for name in golden_names():
        rec = load_golden(name)
        for t in range(int(rec["T"])):
            ues, counts = np.unique(rec["rbg_to_ue"][t][rec["rbg_to_ue"][t] >= 0], return_counts=True)
            nrb = counts * int(rec["rbg_size"])
            hit += int(((nrb > 110) & (nrb % 5 == 0)).sum())
    assert hit > 50, hit
