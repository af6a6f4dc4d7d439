"""Long-horizon parity: >= 1000 TTIs of the unmodified reference on the headline shape and the full
12-s run of BASELINE configs[0], kept as per-100-TTI SHA-256 digests + state checkpoints
(tests/golden/long/, tools/make_golden_long.py).  CPU: the oracle reproduces them; GPU: the CUDA path does."""
import numpy as np
import pytest

from tests.long_replay import load_long, long_names, replay_long

NAMES = long_names()


def _sched(kind, rec):
    S, U = int(rec["S"]), int(rec["U"])
    if kind == "oracle":
        from oracle.pyoracle import OracleScheduler
        return OracleScheduler(int(rec["algo"]), rec["weight"], rec["params"], rec["ue_to_slice"], 1)
    from radiosaber_b200 import sched
    return sched.Scheduler(int(rec["algo"]), rec["weight"], rec["params"], rec["ue_to_slice"], 1)


def test_long_records_exist():
    assert "a9_fix20x5_synth_long" in NAMES and "a9_diffw_trace_12s" in NAMES
    r = load_long("a9_fix20x5_synth_long")
    assert int(r["T"]) >= 1000 and int(r["S"]) == 20 and int(r["U"]) == 100
    r = load_long("a9_diffw_trace_12s")
    assert int(r["T"]) >= 11900 and int(r["U"]) == 204 and len(r["slab_start"]) >= 290   # a CQI report every 40 ms


@pytest.mark.parametrize("name", NAMES)
def test_oracle_reproduces_the_reference_over_the_whole_run(name):
    rec = load_long(name)
    o = _sched("oracle", rec)
    assert replay_long(o, rec) == []          # every block (120 of them for the 12-s run), chained on the oracle's own state


@pytest.mark.gpu
@pytest.mark.parametrize("name", NAMES)
def test_cuda_reproduces_the_reference_over_the_whole_run(name):
    rec = load_long(name)
    g = _sched("cuda", rec)
    assert replay_long(g, rec) == []          # every block, chained on the device's own state
    g.close()


@pytest.mark.gpu
def test_cuda_block_restart_from_checkpoints():
    """Any block replays from the previous block's checkpoint (what a resumed batch run does)."""
    rec = load_long("a9_fix20x5_synth_long")
    g = _sched("cuda", rec)
    assert replay_long(g, rec, blocks=[12, 3, 7]) == []
    g.close()
