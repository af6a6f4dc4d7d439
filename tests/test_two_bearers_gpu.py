"""GPU: two bearers per UE in BATCH mode: per-bearer state on the device
([B][U][2] rates, byte / RB counters, queues, head-of-line delays), slice_priority_ computed per TTI, the priority
hand-over of DoStopSchedule.  Golden = per-bearer records of the unmodified reference on
tests/data/cfg_two_bearers.json (tools/make_golden_two_bearers.py), replayed as cell 3 of a batch whose other cells
carry shuffled queues."""
import os

import numpy as np
import pytest

from radiosaber_b200 import sched, workload
from tests.helpers import ROOT
from tools import golden_io

pytestmark = pytest.mark.gpu
DIR = os.path.join(ROOT, "tests", "golden", "two_bearers")


@pytest.mark.parametrize("algo", [9, 8, 7, 10, 101, 103, 11])
@pytest.mark.parametrize("mode", ["step", "run_host"])
def test_batch_path_reproduces_the_reference_with_two_bearers(algo, mode):
    rec = golden_io.load_npz(os.path.join(DIR, f"a{algo}.npz"))
    T, U, S, G, seed = int(rec["T"]), int(rec["U"]), int(rec["S"]), int(rec["G"]), int(rec["seed"])
    B, CELL = 5, 3
    g = sched.Scheduler(algo, rec["weight"], rec["params"], rec["ue_to_slice"], B, n_bearers=2)
    rng = np.random.default_rng(algo)
    cqi1 = workload.synth_cqi(seed, 0, 1, 0, T, U, G)[:, 0]
    cqi = np.repeat(cqi1[:, None], B, axis=1)
    r2 = np.repeat(rec["rand2"][:, None], B, axis=1)
    if algo == 11:   # every rand() value the reference's 300-sample search drew (cell 3 has the reference's listed users)
        r2 = np.repeat(rec["rand_ng"][:, None], B, axis=1)
    queue = np.stack([rec["queue"][rng.permutation(T)] if b != CELL else rec["queue"] for b in range(B)], axis=1)
    hol = np.repeat(rec["hol"][:, None], B, axis=1)
    avg0 = np.repeat(rec["avg_before"][0][None], B, axis=0)
    tx0 = np.repeat(rec["tx_before"][0][None], B, axis=0)
    st0 = np.repeat(rec["state_before"][0][None], B, axis=0)
    transport = algo in (8, 9, 10, 101, 103)
    g.set_state(avg_rate=avg0, tx_bytes=tx0, slice_offset=st0 if transport else None, nvs_ewma=None if transport else st0)
    ex = rec["exists"].astype(bool)

    def check_state(t):
        st = g.get_state()
        for k, gk in (("avg_rate", "avg_after"), ("tx_bytes", "tx_after"), ("cum_bytes", "cum_bytes"), ("cum_rbs", "cum_rbs")):
            assert np.array_equal(st[k][CELL][ex], rec[gk][t][ex]), (t, k)
        key = "slice_offset" if transport else "nvs_ewma"
        assert np.array_equal(st[key][CELL], rec["state_after"][t]), (t, key)

    def check_out(out, t, i):
        assert np.array_equal(out["tbs_bits"][i][CELL], rec["bits"][t]), (t, "bits")
        assert np.array_equal(out["final_cqi"][i][CELL], rec["final_cqi"][t]), (t, "final_cqi")
        if algo == 10:
            n, m = int(out["alloc_n"][i][CELL]), int(rec["alloc_n"][t])
            assert n == m, (t, "alloc_n")
            from tests.helpers import grants_by_user
            assert grants_by_user(out["alloc_ue"][i][CELL][:n], out["alloc_rbg"][i][CELL][:n]) == \
                grants_by_user(rec["alloc_ue"][t][:m], rec["alloc_rbg"][t][:m]), (t, "grants")
        else:
            assert np.array_equal(out["rbg_to_ue"][i][CELL], rec["rbg_to_ue"][t]), (t, "rbg_to_ue")
        if transport:
            assert np.array_equal(out["slice_target"][i][CELL], rec["target"][t]), (t, "target")
            assert np.array_equal(out["slice_quota"][i][CELL], rec["quota"][t]), (t, "quota")
        else:
            assert int(out["nvs_slice"][i][CELL]) == int(rec["nvs_slice"][t]), (t, "nvs_slice")

    if mode == "step":
        for t in range(T):
            out = g.step(cqi[t], r2[t], dt=float(rec["dt"][t]), want_aux=True, queue=queue[t], hol=hol[t])
            check_out({k: v[None] for k, v in out.items()}, t, 0)
            check_state(t)
    else:
        for t0 in range(0, T, 50):
            out = g.run_host(cqi[t0:t0 + 50], r2[t0:t0 + 50], rec["dt"][t0:t0 + 50], want_aux=True, ttis_per_launch=16,
                             queue=queue[t0:t0 + 50], hol=hol[t0:t0 + 50])
            for i in range(50):
                check_out(out, t0 + i, i)
            check_state(t0 + 49)
    stats = g.get_stats()
    st = g.get_state()
    assert int(stats[0].sum()) == int(st["cum_bytes"].sum())      # the per-slice totals count both bearers
    g.close()


def test_id1_schedules_flows_one_user_per_bearer():
    """DL_PF_PacketScheduler lists FLOWS (one per bearer, downlink-packet-scheduler.cpp:48-94), so a cell with two
    bearers per UE is a cell of one "user" per bearer for id 1: flow f = the f-th bearer of the eNB's container, with
    its UE's CQI row.  Per-bearer rates and counters of the unmodified reference, TTI by TTI."""
    rec = golden_io.load_npz(os.path.join(DIR, "a1.npz"))
    T, U, S, G, seed = int(rec["T"]), int(rec["U"]), int(rec["S"]), int(rec["G"]), int(rec["seed"])
    app = rec["app_id"]
    F = int(app.max()) + 1
    fu = np.zeros(F, np.int32)
    fi = np.zeros(F, np.int32)
    for u in range(U):
        for i in range(2):
            if app[u, i] >= 0:
                fu[app[u, i]], fi[app[u, i]] = u, i
    g = sched.Scheduler(1, rec["weight"], rec["params"], rec["ue_to_slice"][fu], 2)
    cqi = workload.synth_cqi(seed, 0, 1, 0, T, U, G)[:, 0]
    g.set_state(avg_rate=np.repeat(rec["avg_before"][0][fu, fi][None], 2, axis=0),
                tx_bytes=np.repeat(rec["tx_before"][0][fu, fi][None], 2, axis=0))
    served = 0
    for t in range(T):
        q = np.repeat(rec["queue"][t][fu, fi][None], 2, axis=0)
        out = g.step(np.repeat(cqi[t][fu][None], 2, axis=0), None, dt=float(rec["dt"][t]), queue=q)
        st = g.get_state()
        for k, gk in (("avg_rate", "avg_after"), ("tx_bytes", "tx_after"), ("cum_bytes", "cum_bytes"), ("cum_rbs", "cum_rbs")):
            assert np.array_equal(st[k][1], rec[gk][t][fu, fi]), (t, k)
        flows = out["rbg_to_ue"][1]
        assert np.array_equal(np.where(flows >= 0, fu[np.maximum(flows, 0)], -1), rec["rbg_to_ue"][t]), (t, "rbg_to_ue")
        served += int((flows >= 0).sum())
    assert served > 1000
    g.close()


def test_two_bearers_need_queues_and_refuse_flow_level_ids():
    rec = golden_io.load_npz(os.path.join(DIR, "a9.npz"))
    for algo in (1,):
        with pytest.raises(sched.RsError, match="two bearers"):
            sched.Scheduler(algo, rec["weight"], rec["params"], rec["ue_to_slice"], 1, n_bearers=2)
    g = sched.Scheduler(9, rec["weight"], rec["params"], rec["ue_to_slice"], 1, n_bearers=2)
    U = int(rec["U"])
    with pytest.raises(sched.RsError, match="rs_set_queues"):
        g.step(workload.synth_cqi(1, 0, 1, 0, 1, U, 64)[0], np.zeros((1, 2), np.int32))
    g.close()


@pytest.mark.parametrize("algo", [9, 8, 7, 10, 101, 103, 11])
def test_two_bearers_random_cells_against_the_oracle(algo):
    """Beyond the reference's records: random slice parameters, queues and delays, two bearers per UE, 30 TTIs of 12
    cells, CUDA against the oracle (outputs every TTI, the per-bearer state at the end)."""
    from oracle.pyoracle import OracleScheduler
    rng = np.random.default_rng(2000 + algo)
    S, B, T = 7, 12, 30
    ues = rng.integers(1, 7, S)
    u2s = np.repeat(np.arange(S), ues).astype(np.int32)
    U, G = len(u2s), 64
    w = rng.dirichlet(np.ones(S))
    p = np.zeros((S, 4), dtype=np.int32)
    p[:, 0] = rng.integers(0, 2, S)
    p[:, 1] = rng.integers(0, 2, S)
    p[:, 2] = rng.choice([0, 1, 1, 2], S)
    p[:, 3] = rng.integers(0, 2, S)
    g = sched.Scheduler(algo, w, p, u2s, B, n_bearers=2)
    o = OracleScheduler(algo, w, p, u2s, B, n_bearers=2, n_threads=4)
    _, dts = workload.tti_clock(T)
    for t in range(T):
        cqi = workload.synth_cqi(algo, 0, B, t, 1, U, G)[0]
        draws = workload.synth_rand_draws(algo, 0, B, t, 1, S, max(g.rand_stride, 2))[0]
        kind = rng.random((B, U, 2))
        queue = np.where(kind < 0.35, 0, np.where(kind < 0.8, rng.integers(20, 5000, (B, U, 2)), 100000000)).astype(np.int32)
        hol = np.where(rng.random((B, U, 2)) < 0.1, 0.0, rng.random((B, U, 2)) * 0.05)
        act = (rng.random((B, U)) < 0.9).astype(np.uint8)
        a = o.step(cqi, draws, dt=float(dts[t]), active=act, want_aux=True, queue=queue, hol=hol)
        b = g.step(cqi, draws, dt=float(dts[t]), active=act, want_aux=True, queue=queue, hol=hol)
        for k in b:
            assert np.array_equal(a[k], b[k]), (t, k)
    sa, sb = o.get_state(), g.get_state()
    for k in ("avg_rate", "tx_bytes", "cum_bytes", "cum_rbs"):
        assert np.array_equal(sa[k], sb[k]), k
    g.close()
