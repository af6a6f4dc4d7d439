"""GPU: the CUDA path, called through the C ABI, against (a) the golden vectors recorded from the
unmodified reference and (b) the CPU oracle on the same seeded inputs.  Bit-exact everywhere:
assignments, TBS, MCS, targets/quotas are integers, and the EWMA rates / offsets / credits are
IEEE doubles produced by the same operations in the same order (tolerance 0, which is inside the
1e-9 relative bound BASELINE.json states)."""
import numpy as np
import pytest

from oracle import pyoracle
from oracle.pyoracle import OracleScheduler
from radiosaber_b200 import sched, workload
from tests.helpers import golden_names, load_golden, replay_golden

pytestmark = pytest.mark.gpu


def _cqi_keys(rng, shape):
    p = workload.CQI_HIST.astype(np.float64) / workload.CQI_TOTAL
    return rng.choice(np.arange(1, 16), size=shape, p=p).astype(np.uint8)


# ---- the std::sort emulation on its own ---------------------------------------------------------
@pytest.mark.parametrize("n", [1, 2, 16, 17, 31, 32, 33, 64, 320, 448, 1280, 1290, 3200, 4096])
def test_device_sort_matches_std_sort(n):
    rng = np.random.default_rng(n)
    rows = [
        _cqi_keys(rng, n),
        rng.integers(0, 16, size=n).astype(np.uint8),
        np.full(n, 7, np.uint8),
        np.sort(rng.integers(0, 16, size=n)).astype(np.uint8),
        np.sort(rng.integers(0, 16, size=n))[::-1].astype(np.uint8),
        np.maximum.reduce(_cqi_keys(rng, (5, n))),
        (rng.integers(0, 2, size=n) * 15).astype(np.uint8),
        np.maximum.reduce(_cqi_keys(rng, (40, n))),
    ]
    for _ in range(24):
        rows.append(np.maximum.reduce(_cqi_keys(rng, (int(rng.integers(1, 8)), n))))
    keys = np.stack(rows)
    got = sched.test_sort(keys)
    for i, row in enumerate(keys):
        want = pyoracle.std_sort_desc(row.astype(np.float64))
        assert np.array_equal(got[i], want), (n, i)


def test_device_sort_equal_key_ranges():
    """Ranges of equal keys take the precomputed-permutation shortcut: every length, and mixtures
    where long equal runs appear only after a few partitions."""
    rng = np.random.default_rng(99)
    for n in list(range(17, 140)) + [255, 256, 257, 511, 777, 1024, 1280, 2047, 2048, 2049, 3000, 4096]:
        keys = np.stack([np.full(n, 9, np.uint8), np.full(n, 0, np.uint8), np.full(n, 15, np.uint8),
                         rng.choice(np.array([3, 12], dtype=np.uint8), size=n),
                         rng.choice(np.array([3, 12, 13], dtype=np.uint8), size=n, p=[0.1, 0.8, 0.1])])
        got = sched.test_sort(keys)
        for i in range(keys.shape[0]):
            want = pyoracle.std_sort_desc(keys[i].astype(np.float64))
            assert np.array_equal(got[i], want), (n, i)
    # shallow depth limits: the shortcut must step aside when the heap-sort fallback would be reached
    for depth in (1, 2, 3, 5):
        for n in (40, 300, 1280):
            keys = np.stack([np.full(n, 5, np.uint8), rng.choice(np.array([3, 12], dtype=np.uint8), size=n)])
            got = sched.test_sort(keys, depth_limit=depth)
            for i in range(keys.shape[0]):
                want = pyoracle.introsort_emul_desc(keys[i].astype(np.float64), depth)
                assert np.array_equal(got[i], want), (n, depth, i)


@pytest.mark.parametrize("depth", [0, 1, 2, 4])
def test_device_sort_heap_fallback(depth):
    rng = np.random.default_rng(50 + depth)
    for n in (17, 100, 1280):
        keys = rng.integers(0, 16, size=(6, n)).astype(np.uint8)
        got = sched.test_sort(keys, depth_limit=depth)
        for i in range(keys.shape[0]):
            want = pyoracle.introsort_emul_desc(keys[i].astype(np.float64), depth)
            assert np.array_equal(got[i], want), (n, depth, i)


# ---- golden vectors from the unmodified reference ----------------------------------------------
@pytest.mark.parametrize("name", golden_names())
def test_cuda_matches_reference_record(name):
    rec = load_golden(name)
    s = sched.Scheduler(int(rec["algo"]), rec["weight"], rec["params"], rec["ue_to_slice"], 1,
                        n_rbs=int(rec["R"]), rbg_size=int(rec["rbg_size"]), cqi_per_rb=int(rec["cqi_per_rb"]))
    bad = replay_golden(s, rec)
    s.close()
    assert not bad, bad[:10]


# ---- batches against the oracle -----------------------------------------------------------------
def _mk(algo, S, ues_per_slice, weights, params, B, seed, T, cqi_per_rb=0, with_active=False, refresh=1,
        with_queue=False):
    rng = np.random.default_rng(seed)
    u2s = np.repeat(np.arange(S), ues_per_slice).astype(np.int32)
    U = len(u2s)
    G, R = 64, 512
    o = OracleScheduler(algo, weights, params, u2s, B, cqi_per_rb=cqi_per_rb, n_threads=8)
    g = sched.Scheduler(algo, weights, params, u2s, B, cqi_per_rb=cqi_per_rb)
    _, dts = workload.tti_clock(T)
    for t in range(T):
        cqi = workload.synth_cqi(seed, 0, B, t, 1, U, G, refresh)[0]
        if cqi_per_rb:
            cqi = np.repeat(cqi, 8, axis=-1)
            noise = rng.integers(-1, 2, size=cqi.shape)
            cqi = np.clip(cqi.astype(np.int64) + noise, 1, 15).astype(np.uint8)
        rand2 = workload.synth_rand_draws(seed, 0, B, t, 1, S, max(g.rand_stride, 2))[0]
        act = None
        if with_active:
            act = (rng.random((B, U)) < 0.7).astype(np.uint8)
            act[0] = 0                     # a cell with nobody to schedule
            act[1, u2s != 1] = 0           # a cell where only slice 1 has data
        kw = {}
        if with_queue:   # finite queues small enough for the id 7 guard and the id 1 cut-off to bind, idle bearers,
            kind = rng.random((B, U))   # infinite buffers; head-of-line delays with exact zeros
            queue = np.where(kind < 0.2, 0, np.where(kind < 0.6, rng.integers(40, 4000, (B, U)), 100000000))
            hol = np.where(rng.random((B, U)) < 0.1, 0.0, rng.random((B, U)) * 0.05)
            kw = {"queue": queue.astype(np.int32), "hol": hol}
        a = o.step(cqi, rand2, dt=float(dts[t]), active=act, want_aux=True, **kw)
        b = g.step(cqi, rand2, dt=float(dts[t]), active=act, want_aux=True, **kw)
        for k in b:
            assert np.array_equal(a[k], b[k]), (t, k, np.argwhere(a[k] != b[k])[:5])
        sa, sb = o.get_state(), g.get_state()
        for k in ("avg_rate", "tx_bytes", "cum_bytes", "cum_rbs"):
            assert np.array_equal(sa[k], sb[k]), (t, k)
        if algo in (8, 9, 10, 101, 103):
            assert np.array_equal(sa["slice_offset"], sb["slice_offset"]), t
        if algo in (7, 11):
            assert np.array_equal(sa["nvs_ewma"], sb["nvs_ewma"]), t
    g.close()


PF = [0, 0, 1, 1]
MT = [0, 0, 1, 0]


@pytest.mark.parametrize("algo", [9, 8, 7, 1, 11, 10, 101, 103])
def test_headline_shape_20x5(algo):
    S = 20
    _mk(algo, S, [5] * S, np.full(S, 0.05), np.tile(PF, (S, 1)), B=48, seed=algo, T=25)


@pytest.mark.parametrize("algo", [10, 11, 101, 103])
def test_f4_ids_steady_state_on_the_headline_cell(algo):
    """The f4 schedulers far past the start-up transient (every bearer starts at the same average rate, so the first TTIs
    are full of ties): 300 TTIs of the headline cell on the compile-time-shape kernels against the oracle, state
    compared after every TTI.  Pins the incremental re-scans of id 103 (three-maxima rule included) and the bitmask search
    of id 11 where the bench-shaped sweeps time them."""
    S = 20
    _mk(algo, S, [5] * S, np.full(S, 0.05), np.tile(PF, (S, 1)), B=8, seed=500 + algo, T=300)


@pytest.mark.parametrize("layout", [0, 1])
def test_nvs_nongreedy_slice_sizes_around_the_bitmask_search(layout):
    """Id 11 scores its 300 samples on RBG bitmasks when the served slice lists at most 8 users (two instantiations: up
    to 5, up to 8) and entry by entry otherwise; inactive bearers move a slice between the three.  NVS serves one slice
    per TTI, so 60 TTIs visit every size."""
    ups = [1, 2, 3, 4, 5, 6, 7, 8, 9, 12, 5, 8]
    S = len(ups)
    _mk(11, S, ups, np.full(S, 1.0 / S), np.tile(PF, (S, 1)), B=12, seed=1100 + layout, T=60, cqi_per_rb=layout, with_active=True)


@pytest.mark.parametrize("algo", [9, 8, 7, 10, 101, 103])
def test_mixed_enterprise_schedulers_diff_weights(algo):
    S = 20
    w = np.array([0.025] * 10 + [0.075] * 10)
    p = np.array([MT] * 10 + [PF] * 10, dtype=np.int32)
    ups = [5, 6, 6, 10, 7, 15, 9, 9, 14, 8, 14, 5, 14, 15, 7, 11, 15, 11, 13, 10]
    _mk(algo, S, ups, w, p, B=16, seed=20 + algo, T=20)


@pytest.mark.parametrize("S,n", [(5, 2), (5, 40), (10, 15), (30, 10), (50, 2), (50, 40), (64, 3), (1, 7), (3, 1)])
def test_sweep_shapes_radiosaber(S, n):
    w = 1.0 + (np.arange(S) % 3)
    w = w / w.sum()
    p = np.array([PF if s % 2 == 0 else MT for s in range(S)], dtype=np.int32)
    _mk(9, S, [n] * S, w, p, B=6, seed=S * 100 + n, T=12)


@pytest.mark.parametrize("algo", [9, 8, 7, 1, 11, 10])
@pytest.mark.parametrize("layout", [0, 1])
def test_wide_cells_all_ids(algo, layout):
    """Cells with >= 480 UEs run 512 threads wide (the second instantiation of the device code)."""
    S = 12
    w = np.array([0.2, 0.1, 0.1, 0.05, 0.05, 0.1, 0.1, 0.1, 0.05, 0.05, 0.05, 0.05])
    p = np.tile(PF, (S, 1))
    p[1::3] = MT
    g = sched.Scheduler(algo, w, p, np.repeat(np.arange(S), 40).astype(np.int32), 1)
    assert sched.lib().rs_threads_per_cta(g._h) == 512
    g.close()
    _mk(algo, S, [40] * S, w, p, B=4, seed=700 + algo, T=5, cqi_per_rb=layout, with_active=(layout == 1))


@pytest.mark.parametrize("algo", [9, 8, 7, 1, 11, 10, 101, 103])
@pytest.mark.parametrize("layout", [0, 1])
def test_queue_aware_enterprise_schedulers(algo, layout):
    """per-TTI queue sizes and head-of-line delays; slices with alpha/beta set (HoL-weighted
    metrics), finite queues (bytes capped, id 7's required-RBs guard, id 1's flow-satisfied cut-off)."""
    S = 6
    w = np.array([0.3, 0.1, 0.2, 0.1, 0.2, 0.1])
    p = np.array([[0, 0, 1, 1], [1, 0, 1, 1], [1, 1, 1, 1], [1, 1, 1, 0], [0, 0, 1, 0], [1, 1, 2, 1]], dtype=np.int32)
    _mk(algo, S, [4, 3, 5, 2, 4, 3], w, p, B=24, seed=900 + algo, T=14, cqi_per_rb=layout, with_queue=True,
        with_active=(layout == 1))


@pytest.mark.parametrize("S,n", [(14, 3), (20, 2), (31, 2), (40, 2), (61, 1), (64, 2)])
def test_subopt_many_slices_hashtable_order(S, n):
    """SubOpt's tie-break is std::unordered_map iteration order: cross the 13 / 29 / 59 bucket counts."""
    rng = np.random.default_rng(S)
    w = rng.dirichlet(np.ones(S) * 0.4)
    _mk(101, S, [n] * S, w, np.tile(PF, (S, 1)), B=12, seed=1200 + S, T=10, with_active=True)


@pytest.mark.parametrize("algo", [9, 8, 7, 1])
def test_queue_aware_wide_cells(algo):
    S = 12
    w = np.full(S, 1.0 / S)
    p = np.tile(np.array([1, 1, 1, 1], dtype=np.int32), (S, 1))
    p[::2] = [0, 0, 1, 1]
    _mk(algo, S, [40] * S, w, p, B=3, seed=950 + algo, T=4, with_queue=True)


@pytest.mark.parametrize("algo", [9, 8, 7, 1, 11, 10, 101, 103])
def test_inactive_bearers_and_empty_cells(algo):
    S = 8
    w = np.full(S, 1.0 / S)
    _mk(algo, S, [4] * S, w, np.tile(PF, (S, 1)), B=10, seed=300 + algo, T=12, with_active=True)


@pytest.mark.parametrize("algo", [9, 8, 7, 1, 11, 10, 101, 103])
def test_per_rb_cqi_layout(algo):
    S = 6
    w = np.full(S, 1.0 / S)
    _mk(algo, S, [3] * S, w, np.tile(PF, (S, 1)), B=6, seed=400 + algo, T=10, cqi_per_rb=1)


def test_epsilon_other_than_one():
    S = 4
    p = np.array([[0, 0, 2, 1], [0, 0, 0, 1], [0, 0, 3, 0], [0, 0, 1, 1]], dtype=np.int32)
    _mk(9, S, [6] * S, np.full(S, 0.25), p, B=8, seed=77, T=10)


def test_cqi_held_for_40_ttis():
    S = 20
    _mk(9, S, [5] * S, np.full(S, 0.05), np.tile(PF, (S, 1)), B=8, seed=5, T=45, refresh=40)


# ---- multi-TTI entry points, generators, statistics ---------------------------------------------
def test_run_host_equals_stepwise_and_oracle():
    S, B, T = 20, 32, 24
    u2s = np.repeat(np.arange(S), 5).astype(np.int32)
    U, G = len(u2s), 64
    w, p = np.full(S, 0.05), np.tile(PF, (S, 1))
    cqi = workload.synth_cqi(9, 0, B, 0, T, U, G)
    rand2 = workload.synth_rand2(9, 0, B, 0, T, S)
    _, dts = workload.tti_clock(T)
    o = OracleScheduler(9, w, p, u2s, B, n_threads=8)
    want = [o.step(cqi[t], rand2[t], dt=float(dts[t])) for t in range(T)]
    for tpl in (1, 5, 24):
        g = sched.Scheduler(9, w, p, u2s, B)
        got = g.run_host(cqi, rand2, dts, ttis_per_launch=tpl)
        for t in range(T):
            for k in ("rbg_to_ue", "tbs_bits", "mcs"):
                assert np.array_equal(got[k][t], want[t][k]), (tpl, t, k)
        st, so = g.get_state(), o.get_state()
        for k in ("avg_rate", "tx_bytes", "cum_bytes", "cum_rbs", "slice_offset"):
            assert np.array_equal(st[k], so[k]), (tpl, k)
        stats = g.get_stats()
        cb = so["cum_bytes"].reshape(B, S, 5)
        assert np.array_equal(stats[0], cb.sum(axis=(0, 2)))
        assert np.array_equal(stats[1], so["cum_rbs"].reshape(B, S, 5).sum(axis=(0, 2)))
        q = cb >> np.uint64(10)
        assert np.array_equal(stats[2], q.sum(axis=(0, 2)))
        assert np.array_equal(stats[3], (q * q).sum(axis=(0, 2)))
        g.close()


def test_device_generators_and_run_device():
    import torch
    S, B, T = 20, 40, 12
    u2s = np.repeat(np.arange(S), 5).astype(np.int32)
    U, G = len(u2s), 64
    w, p = np.full(S, 0.05), np.tile(PF, (S, 1))
    g = sched.Scheduler(9, w, p, u2s, B)
    d_cqi = torch.empty((T, B, U, G), dtype=torch.uint8, device="cuda")
    d_r2 = torch.empty((T, B, 2), dtype=torch.int32, device="cuda")
    g.synth_cqi(3, 100, 7, T, d_cqi.data_ptr())
    g.synth_rand2(3, 100, 7, T, d_r2.data_ptr())
    g.sync()
    cqi = workload.synth_cqi(3, 100, B, 7, T, U, G)
    rand2 = workload.synth_rand2(3, 100, B, 7, T, S)
    assert np.array_equal(d_cqi.cpu().numpy(), cqi)
    assert np.array_equal(d_r2.cpu().numpy(), rand2)
    d_rbg = torch.empty((T, B, G), dtype=torch.int16, device="cuda")
    d_bits = torch.empty((T, B, U), dtype=torch.int32, device="cuda")
    d_mcs = torch.empty((T, B, U), dtype=torch.uint8, device="cuda")
    _, dts = workload.tti_clock(T)
    g.run_device(T, d_cqi.data_ptr(), B * U * G, d_r2.data_ptr(), dts,
                 {"rbg_to_ue": d_rbg.data_ptr(), "tbs_bits": d_bits.data_ptr(), "mcs": d_mcs.data_ptr()},
                 ttis_per_launch=5)
    g.sync()
    o = OracleScheduler(9, w, p, u2s, B, n_threads=8)
    for t in range(T):
        want = o.step(cqi[t], rand2[t], dt=float(dts[t]))
        assert np.array_equal(d_rbg[t].cpu().numpy(), want["rbg_to_ue"]), t
        assert np.array_equal(d_bits[t].cpu().numpy(), want["tbs_bits"]), t
        assert np.array_equal(d_mcs[t].cpu().numpy(), want["mcs"]), t
    assert g.launch_count == 2 + 3
    g.close()


def test_nvs_nongreedy_device_draws_and_run_device():
    """id 11: the device generator of the 300 x users rand() draws equals its numpy twin, and rs_run_device
    over several launches equals the oracle fed with the same draws."""
    import torch
    S, B, T = 6, 33, 9
    ues = [5, 3, 7, 1, 4, 2]
    u2s = np.repeat(np.arange(S), ues).astype(np.int32)
    U, G = len(u2s), 64
    w = np.array([0.3, 0.1, 0.2, 0.1, 0.2, 0.1])
    p = np.tile(PF, (S, 1))
    g = sched.Scheduler(11, w, p, u2s, B, cqi_per_rb=2)
    n = g.rand_stride
    assert n == 300 * 7
    d_cqi = torch.empty((T, B, U, G // 2), dtype=torch.uint8, device="cuda")
    d_r = torch.empty((T, B, n), dtype=torch.int32, device="cuda")
    g.synth_cqi(5, 10, 0, T, d_cqi.data_ptr())
    g.synth_rand2(5, 10, 0, T, d_r.data_ptr())
    g.sync()
    draws = workload.synth_rand_draws(5, 10, B, 0, T, S, n)
    assert np.array_equal(d_r.cpu().numpy(), draws)
    cqi = workload.synth_cqi(5, 10, B, 0, T, U, G)
    d_rbg = torch.empty((T, B, G), dtype=torch.int16, device="cuda")
    d_bits = torch.empty((T, B, U), dtype=torch.int32, device="cuda")
    d_nvs = torch.empty((T, B), dtype=torch.int32, device="cuda")
    _, dts = workload.tti_clock(T)
    g.run_device(T, d_cqi.data_ptr(), B * U * G // 2, d_r.data_ptr(), dts,
                 {"rbg_to_ue": d_rbg.data_ptr(), "tbs_bits": d_bits.data_ptr(), "nvs_slice": d_nvs.data_ptr()},
                 ttis_per_launch=4)
    g.sync()
    o = OracleScheduler(11, w, p, u2s, B, n_threads=8)
    for t in range(T):
        want = o.step(cqi[t], draws[t], dt=float(dts[t]), want_aux=True)
        assert np.array_equal(d_nvs[t].cpu().numpy(), want["nvs_slice"]), t
        assert np.array_equal(d_rbg[t].cpu().numpy(), want["rbg_to_ue"]), t
        assert np.array_equal(d_bits[t].cpu().numpy(), want["tbs_bits"]), t
    assert np.array_equal(g.get_state()["avg_rate"], o.get_state()["avg_rate"])
    g.close()


@pytest.mark.parametrize("algo", [9, 8, 7, 1])
def test_packed_cqi_layout_and_refresh(algo):
    """4-bit CQI layout (two RBGs per byte) and CQI held for several TTIs, through rs_run_host and
    rs_run_device with the device generator: same results as the oracle on the unpacked values."""
    import torch
    S, B, T, refresh = 20, 24, 23, 5
    u2s = np.repeat(np.arange(S), 5).astype(np.int32)
    U, G = len(u2s), 64
    w, p = np.full(S, 0.05), np.tile(PF, (S, 1))
    n_slabs = -(-T // refresh)
    slabs = workload.synth_cqi(11, 50, B, 0, n_slabs, U, G)      # epoch e == "tti" e of the generator
    rand2 = workload.synth_rand2(11, 50, B, 0, T, S)
    _, dts = workload.tti_clock(T)
    o = OracleScheduler(algo, w, p, u2s, B, n_threads=8)
    want = [o.step(slabs[t // refresh], rand2[t], dt=float(dts[t])) for t in range(T)]
    # host path, packed
    g = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=2)
    got = g.run_host(sched.pack_cqi(slabs), rand2, dts, ttis_per_launch=4, cqi_refresh=refresh)
    for t in range(T):
        for k in ("rbg_to_ue", "tbs_bits", "mcs"):
            assert np.array_equal(got[k][t], want[t][k]), (t, k)
    assert np.array_equal(g.get_state()["avg_rate"], o.get_state()["avg_rate"])
    # device path: generator writes the packed layout directly
    g.reset_state()
    d_cqi = torch.empty((n_slabs, B, U, G // 2), dtype=torch.uint8, device="cuda")
    d_r2 = torch.from_numpy(rand2).cuda()
    g.synth_cqi(11, 50, 0, n_slabs, d_cqi.data_ptr())
    g.sync()
    assert np.array_equal(d_cqi.cpu().numpy(), sched.pack_cqi(slabs))
    d_rbg = torch.empty((T, B, G), dtype=torch.int16, device="cuda")
    g.run_device(T, d_cqi.data_ptr(), B * U * G // 2, d_r2.data_ptr(), dts, {"rbg_to_ue": d_rbg.data_ptr()},
                 ttis_per_launch=7, cqi_refresh=refresh)
    g.sync()
    for t in range(T):
        assert np.array_equal(d_rbg[t].cpu().numpy(), want[t]["rbg_to_ue"]), t
    g.close()


def test_full_size_batch_invariants():
    """BASELINE config #2 size (4096 cells x 20 slices x 5 UEs): properties that need no oracle run,
    plus an oracle spot check on a few cells."""
    import torch
    S, B, T = 20, 4096, 6
    u2s = np.repeat(np.arange(S), 5).astype(np.int32)
    U, G = len(u2s), 64
    w, p = np.full(S, 0.05), np.tile(PF, (S, 1))
    g = sched.Scheduler(9, w, p, u2s, B)
    d_cqi = torch.empty((T, B, U, G), dtype=torch.uint8, device="cuda")
    d_r2 = torch.empty((T, B, 2), dtype=torch.int32, device="cuda")
    g.synth_cqi(1, 0, 0, T, d_cqi.data_ptr())
    g.synth_rand2(1, 0, 0, T, d_r2.data_ptr())
    d_rbg = torch.empty((T, B, G), dtype=torch.int16, device="cuda")
    d_bits = torch.empty((T, B, U), dtype=torch.int32, device="cuda")
    d_tgt = torch.empty((T, B, S), dtype=torch.int32, device="cuda")
    d_quo = torch.empty((T, B, S), dtype=torch.int32, device="cuda")
    _, dts = workload.tti_clock(T)
    g.run_device(T, d_cqi.data_ptr(), B * U * G, d_r2.data_ptr(), dts,
                 {"rbg_to_ue": d_rbg.data_ptr(), "tbs_bits": d_bits.data_ptr(),
                  "slice_target": d_tgt.data_ptr(), "slice_quota": d_quo.data_ptr()})
    g.sync()
    rbg = d_rbg.cpu().numpy().astype(np.int64)
    bits = d_bits.cpu().numpy()
    tgt, quo = d_tgt.cpu().numpy(), d_quo.cpu().numpy()
    assert (rbg >= 0).all() and (rbg < U).all()                    # every RBG allocated (backlogged cells)
    assert (tgt.sum(-1) == 512).all() and (quo.sum(-1) == 64).all()  # targets / quotas are partitions
    per_slice = np.zeros((T, B, S), dtype=np.int64)
    np.add.at(per_slice, (np.arange(T)[:, None, None], np.arange(B)[None, :, None], u2s[rbg]), 1)
    assert (per_slice <= np.maximum(quo, 0)).all()                 # first-fit never exceeds a quota
    got_ue = np.zeros((T, B, U), dtype=np.int64)
    np.add.at(got_ue, (np.arange(T)[:, None, None], np.arange(B)[None, :, None], rbg), 1)
    assert ((bits > 0) == (got_ue > 0)).all()                      # bits iff RBGs
    st = g.get_state()
    assert np.array_equal(st["cum_rbs"].astype(np.int64), got_ue.sum(0) * 8)
    assert np.array_equal(st["cum_bytes"].astype(np.int64), (bits // 8).astype(np.int64).sum(0))
    # oracle spot check: first and last 8 cells
    cqi = d_cqi.cpu().numpy()
    r2 = d_r2.cpu().numpy()
    for sl in (slice(0, 8), slice(B - 8, B)):
        o = OracleScheduler(9, w, p, u2s, 8, n_threads=8)
        for t in range(T):
            want = o.step(cqi[t, sl], r2[t, sl], dt=float(dts[t]))
            assert np.array_equal(want["rbg_to_ue"], rbg[t, sl]), t
            assert np.array_equal(want["tbs_bits"], bits[t, sl]), t
        assert np.array_equal(o.get_state()["avg_rate"], st["avg_rate"][sl])
    g.close()


@pytest.mark.parametrize("algo", [9, 8, 10, 101, 103, 7, 11])
@pytest.mark.parametrize("layout", [0, 2])
def test_fixed_shape_equals_dynamic(algo, layout):
    """The headline cell runs a compile-time-shape instantiation of the TTI kernel; RS_NO_FIXED_SHAPE keeps a handle on
    the general one.  Same inputs, 70 TTIs in 16-TTI launches (streamed CQI and trace replay): every output and the
    whole state are identical, and both equal the oracle."""
    import os
    from oracle.pyoracle import OracleScheduler
    S, n, B, T = 20, 5, 40, 70
    w = np.full(S, 0.05)
    p = np.tile(PF, (S, 1))
    p[1::3] = MT
    u2s = np.repeat(np.arange(S), n).astype(np.int32)
    U, G = len(u2s), 64
    fixed = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=layout)
    os.environ["RS_NO_FIXED_SHAPE"] = "1"
    try:
        dyn = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=layout)
    finally:
        del os.environ["RS_NO_FIXED_SHAPE"]
    assert sched.lib().rs_fixed_shape(fixed._h) >= 0 and sched.lib().rs_fixed_shape(dyn._h) == -1
    odd = sched.Scheduler(algo, w, p, np.repeat(np.arange(S), n)[::-1].astype(np.int32).copy(), B, cqi_per_rb=layout)
    assert sched.lib().rs_fixed_shape(odd._h) == -1          # same sizes, another UE -> slice map: general kernel
    odd.close()
    cqi = workload.synth_cqi(77, 0, B, 0, T, U, G)
    r2 = workload.synth_rand_draws(77, 0, B, 0, T, S, max(fixed.rand_stride, 2))   # id 11: 300 draws per user of the largest slice
    _, dts = workload.tti_clock(T)
    feed = sched.pack_cqi(cqi) if layout == 2 else cqi
    a = fixed.run_host(feed, r2, dts, want_aux=True, ttis_per_launch=16)
    b = dyn.run_host(feed, r2, dts, want_aux=True, ttis_per_launch=16)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    sa, sb = fixed.get_state(), dyn.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    o = OracleScheduler(algo, w, p, u2s, B, n_threads=8)
    for t in range(T):
        r = o.step(cqi[t], r2[t], dt=float(dts[t]))
        assert np.array_equal(r["rbg_to_ue"], a["rbg_to_ue"][t]) and np.array_equal(r["tbs_bits"], a["tbs_bits"][t]), t
    assert np.array_equal(o.get_state()["avg_rate"], sa["avg_rate"])
    # trace replay through both
    rng = np.random.default_rng(5)
    traces = np.repeat(workload.histogram_cqi(rng, (12, 30, G)), 8, axis=2)
    ue_trace = rng.integers(0, 12, (B, U)).astype(np.int32)
    now, _ = workload.tti_clock(T)
    rows = sched.trace_rows_for_run(now, 0, n_rows=30)
    outs = []
    for g in (fixed, dyn):
        g.reset_state()
        g.set_traces(traces, ue_trace)
        outs.append((g.run_traces_host(rows, r2, dts, want_aux=True, ttis_per_launch=16), g.get_state()))
    for k in outs[0][0]:
        assert np.array_equal(outs[0][0][k], outs[1][0][k]), ("trace", k)
    for k in outs[0][1]:
        assert np.array_equal(outs[0][1][k], outs[1][1][k]), ("trace", k)
    fixed.close()
    dyn.close()


@pytest.mark.parametrize("algo,S,n,queue", [(9, 20, 20, False), (8, 12, 40, False), (9, 12, 40, True), (10, 6, 24, False),
                                            (101, 10, 10, True), (9, 50, 10, False)])
def test_direct_metric_equals_table(algo, S, n, queue):
    """Big slices divide their metrics on the fly (rs_direct_metric); RS_NO_DIRECT keeps the per-chunk table.  Same
    inputs, every output and the state identical (and both equal the oracle: the shapes of test_sweep_shapes_* and
    test_wide_cells_* run the direct path against it)."""
    import os
    B, T = 6, 10
    rng = np.random.default_rng(S * n)
    w = rng.dirichlet(np.ones(S))
    p = np.array([PF if s % 3 else MT for s in range(S)], dtype=np.int32)
    if queue:
        p[1] = [1, 1, 1, 1]
        p[2] = [1, 0, 1, 1]
    u2s = np.repeat(np.arange(S), n).astype(np.int32)
    U, G = len(u2s), 64
    direct = sched.Scheduler(algo, w, p, u2s, B)
    os.environ["RS_NO_DIRECT"] = "1"
    try:
        table = sched.Scheduler(algo, w, p, u2s, B)
    finally:
        del os.environ["RS_NO_DIRECT"]
    assert sched.lib().rs_direct_metric(direct._h) == 1 and sched.lib().rs_direct_metric(table._h) == 0
    _, dts = workload.tti_clock(T)
    for t in range(T):
        cqi = workload.synth_cqi(9, 0, B, t, 1, U, G)[0]
        r2 = workload.synth_rand_draws(9, 0, B, t, 1, S, max(direct.rand_stride, 2))[0]
        kw = {}
        if queue:
            kw = {"queue": rng.choice([0, 500, 100000000], size=(B, U)).astype(np.int32), "hol": rng.random((B, U)) * 0.03}
        act = (rng.random((B, U)) < 0.85).astype(np.uint8)
        a = direct.step(cqi, r2, dt=float(dts[t]), active=act, want_aux=True, **kw)
        b = table.step(cqi, r2, dt=float(dts[t]), active=act, want_aux=True, **kw)
        for k in a:
            assert np.array_equal(a[k], b[k]), (t, k)
    sa, sb = direct.get_state(), table.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    direct.close()
    table.close()
