"""CPU: the oracle with two bearers per UE (rso_config.n_bearers = 2) reproduces the per-bearer records of the unmodified
reference (tests/golden/two_bearers/, tools/make_golden_two_bearers.py): allocations, allocated bits, every bearer's
EWMA rate and byte / RB counters after each of 200 TTIs, ids 9, 8, 7, 10, 101, 103, 11."""
import os

import numpy as np
import pytest

from oracle.pyoracle import OracleScheduler
from radiosaber_b200 import workload
from tests.helpers import ROOT, grants_by_user
from tools import golden_io

DIR = os.path.join(ROOT, "tests", "golden", "two_bearers")


@pytest.mark.parametrize("algo", [9, 8, 7, 10, 101, 103, 11])
def test_oracle_reproduces_the_reference_with_two_bearers(algo):
    rec = golden_io.load_npz(os.path.join(DIR, f"a{algo}.npz"))
    T, U, G, seed = int(rec["T"]), int(rec["U"]), int(rec["G"]), int(rec["seed"])
    o = OracleScheduler(algo, rec["weight"], rec["params"], rec["ue_to_slice"], 1, n_bearers=2)
    transport = algo in (8, 9, 10, 101, 103)
    o.set_state(avg_rate=rec["avg_before"][0][None], tx_bytes=rec["tx_before"][0][None],
                slice_offset=rec["state_before"][0][None] if transport else None,
                nvs_ewma=None if transport else rec["state_before"][0][None])
    cqi = workload.synth_cqi(seed, 0, 1, 0, T, U, G)[:, 0]
    ex = rec["exists"].astype(bool)
    both = 0
    for t in range(T):
        draws = rec["rand_ng"][t][None] if algo == 11 else rec["rand2"][t][None]
        out = o.step(cqi[t][None], draws, dt=float(rec["dt"][t]), want_aux=True, queue=rec["queue"][t][None], hol=rec["hol"][t][None])
        st = o.get_state()
        assert np.array_equal(out["tbs_bits"][0], rec["bits"][t]), t
        assert np.array_equal(out["final_cqi"][0], rec["final_cqi"][t]), t
        if algo == 10:
            n, m = int(out["alloc_n"][0]), int(rec["alloc_n"][t])
            assert n == m and grants_by_user(out["alloc_ue"][0][:n], out["alloc_rbg"][0][:n]) == \
                grants_by_user(rec["alloc_ue"][t][:m], rec["alloc_rbg"][t][:m]), t
        else:
            assert np.array_equal(out["rbg_to_ue"][0], rec["rbg_to_ue"][t]), t
        for k, gk in (("avg_rate", "avg_after"), ("tx_bytes", "tx_after"), ("cum_bytes", "cum_bytes"), ("cum_rbs", "cum_rbs")):
            assert np.array_equal(st[k][0][ex], rec[gk][t][ex]), (t, k)
        assert np.array_equal(st["slice_offset" if transport else "nvs_ewma"][0], rec["state_after"][t]), t
        both += int(((rec["queue"][t][:, 0] > 0) & (rec["queue"][t][:, 1] > 0)).sum())
    assert both > 100          # the records do exercise UEs with both bearers queued


def test_two_bearers_refused_where_the_reference_schedules_flows():
    rec = golden_io.load_npz(os.path.join(DIR, "a9.npz"))
    o = OracleScheduler(1, rec["weight"], rec["params"], rec["ue_to_slice"], 1, n_bearers=2)
    with pytest.raises(RuntimeError):
        o.step(np.full((1, int(rec["U"]), 64), 7, np.uint8), None, queue=rec["queue"][0][None])
