"""GPU: the host side of the C ABI added in round 2 -- rs_step_cell (the one-round-trip TTI of the in-simulator
plug-in), the slot ring that stays in flight across rs_run_host_async calls, two live handles of different cell
sizes (the shared-memory attribute belongs to the kernel function), rs_reduce_stats, rs_batch --gpus."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

from radiosaber_b200 import sched, workload
from tests.helpers import ROOT

pytestmark = pytest.mark.gpu
PF, MT = [0, 0, 1, 1], [0, 0, 1, 0]


def _cfg(S, n, mix=False):
    w = np.full(S, 1.0 / S)
    p = np.array([MT if (mix and s % 2) else PF for s in range(S)], dtype=np.int32)
    return w, p, np.repeat(np.arange(S), n).astype(np.int32)


def test_two_live_handles_of_different_cell_sizes():
    """create the big handle first, then a small one of the same instantiation, then launch the big one."""
    wb, pb, ub = _cfg(20, 20)     # 400 UEs: rs:: instantiation, ~50 KB of shared memory
    ws, ps, us = _cfg(4, 2)
    big = sched.Scheduler(9, wb, pb, ub, 4)
    small = sched.Scheduler(9, ws, ps, us, 4)
    assert big.smem_bytes > small.smem_bytes
    assert sched.lib().rs_threads_per_cta(big._h) == sched.lib().rs_threads_per_cta(small._h)
    for g, (w, p, u) in ((big, (wb, pb, ub)), (small, (ws, ps, us)), (big, (wb, pb, ub))):
        U, S = len(u), len(w)
        out = g.step(workload.synth_cqi(1, 0, 4, 0, 1, U, 64)[0], workload.synth_rand2(1, 0, 4, 0, 1, S)[0])
        assert (out["rbg_to_ue"] >= 0).all()
    big.close()
    small.close()


@pytest.mark.parametrize("algo", [9, 8, 7, 1, 10, 11])
@pytest.mark.parametrize("layout", [0, 1])
def test_step_cell_equals_the_four_call_sequence(algo, layout):
    S, n, B, T = 6, 4, 3, 12
    w, p, u2s = _cfg(S, n, mix=True)
    p[2] = [1, 1, 1, 1]
    p[3] = [1, 0, 1, 1]
    U = len(u2s)
    a = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=layout)
    b = sched.Scheduler(algo, w, p, u2s, B, cqi_per_rb=layout)
    rng = np.random.default_rng(algo * 10 + layout)
    avg = np.full((B, U), 100000.0)
    st = np.zeros((B, S))
    _, dts = workload.tti_clock(T)
    for t in range(T):
        cqi = workload.synth_cqi(3, 0, B, t, 1, U, 64)[0]
        if layout:
            cqi = np.repeat(cqi, 8, axis=-1)
        draws = workload.synth_rand_draws(3, 0, B, t, 1, S, max(a.rand_stride, 2))[0]
        queue = rng.choice([0, 900, 100000000], size=(B, U)).astype(np.int32)
        hol = rng.random((B, U)) * 0.02
        act = (rng.random((B, U)) < 0.8).astype(np.uint8)
        want = a.step(cqi, draws, dt=float(dts[t]), active=act, want_aux=True, queue=queue, hol=hol)
        sa = a.get_state()
        got = b.step_cell(cqi, draws, dt=float(dts[t]), active=act, queue=queue, hol=hol, avg_rate=avg, slice_state=st)
        for k in want:
            assert np.array_equal(want[k], got[k]), (t, k)
        assert np.array_equal(avg, sa["avg_rate"]), t
        if algo in (8, 9, 10):
            assert np.array_equal(st, sa["slice_offset"]), t
        if algo in (7, 11):
            assert np.array_equal(st, sa["nvs_ewma"]), t
    a.close()
    b.close()


@pytest.mark.parametrize("refresh", [1, 40])
def test_async_calls_keep_the_ring_in_flight(refresh):
    """Three rs_run_host_async calls back to back (chunks that do not divide the calls, a CQI slab per `refresh`
    TTIs) equal one synchronous run of the same TTIs."""
    S, n, B = 20, 5, 24
    w, p, u2s = _cfg(S, n)
    U, G = len(u2s), 64
    lens = [50, 33, 40]
    T = sum(lens)
    _, dts = workload.tti_clock(T)
    ref = sched.Scheduler(9, w, p, u2s, B, cqi_per_rb=2)
    g = sched.Scheduler(9, w, p, u2s, B, cqi_per_rb=2)
    r2 = workload.synth_rand2(5, 0, B, 0, T, S)
    outs, keep, tickets = [], [], []
    t0 = 0
    want = {k: [] for k in ("rbg_to_ue", "tbs_bits", "mcs")}
    for L in lens:
        n_slabs = -(-L // refresh)
        cqi = sched.pack_cqi(workload.synth_cqi(5, 0, B, t0, n_slabs, U, G))   # the call's own slabs
        rr = np.ascontiguousarray(r2[t0:t0 + L])
        dd = np.ascontiguousarray(dts[t0:t0 + L])
        a = ref.run_host(cqi, rr, dd, ttis_per_launch=16, cqi_refresh=refresh)
        for k in want:
            want[k].append(a[k])
        out, o = g._host_outputs(L, False)
        keep.append((cqi, rr, dd, out, o))
        tickets.append(g.run_host_async(cqi, rr, dd, out, o, ttis_per_launch=7, cqi_refresh=refresh))
        outs.append(out)
        t0 += L
    g.wait(tickets[0])
    assert np.array_equal(outs[0]["rbg_to_ue"], want["rbg_to_ue"][0])
    g.wait(tickets[2])
    for i in range(3):
        for k in want:
            assert np.array_equal(outs[i][k], want[k][i]), (i, k)
    sa, sb = ref.get_state(), g.get_state()
    for k in ("avg_rate", "cum_bytes", "slice_offset"):
        assert np.array_equal(sa[k], sb[k]), k
    with pytest.raises(sched.RsError, match="never issued"):
        g.wait(99)
    ref.close()
    g.close()


def test_headline_shape_long_run_against_the_oracle():
    """the headline shape, id 9, 480 TTIs of 32 cells through rs_run_host in 16-TTI launches
    against the oracle stepping TTI by TTI -- the steady state the bench times, not its first 25 TTIs."""
    from oracle.pyoracle import OracleScheduler
    S, n, B, T = 20, 5, 32, 480
    w, p, u2s = _cfg(S, n)
    U, G = len(u2s), 64
    g = sched.Scheduler(9, w, p, u2s, B)
    o = OracleScheduler(9, w, p, u2s, B, n_threads=os.cpu_count() or 1)
    _, dts = workload.tti_clock(T)
    for t0 in range(0, T, 96):
        cqi = workload.synth_cqi(11, 0, B, t0, 96, U, G)
        r2 = workload.synth_rand2(11, 0, B, t0, 96, S)
        got = g.run_host(cqi, r2, dts[t0:t0 + 96], ttis_per_launch=16)
        for t in range(96):
            a = o.step(cqi[t], r2[t], dt=float(dts[t0 + t]))
            for k in ("rbg_to_ue", "tbs_bits", "mcs"):
                assert np.array_equal(a[k], got[k][t]), (t0 + t, k)
        sa, sb = o.get_state(), g.get_state()
        for k in ("avg_rate", "tx_bytes", "cum_bytes", "cum_rbs", "slice_offset"):
            assert np.array_equal(sa[k], sb[k]), (t0, k)
    g.close()


def _nccl():
    import torch  # noqa: F401  -- first: librs_nccl.so binds to whichever libnccl.so.2 the process has loaded, and torch
    #                              must get its own bundled one (the system's is older and lacks symbols torch needs)
    L = C.CDLL(os.path.join(ROOT, "radiosaber_b200", "librs_nccl.so"))
    L.rs_nccl_last_error.restype = C.c_char_p
    L.rs_reduce_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.rs_comm_destroy.argtypes = [C.c_void_p]
    return L


def test_reduce_stats_single_rank_communicator():
    w, p, u2s = _cfg(5, 3)
    g = sched.Scheduler(9, w, p, u2s, 16)
    _, dts = workload.tti_clock(20)
    g.run_host(workload.synth_cqi(2, 0, 16, 0, 20, 15, 64), workload.synth_rand2(2, 0, 16, 0, 20, 5), dts)
    L = _nccl()
    uid = (C.c_uint8 * 128)()
    assert L.rs_comm_unique_id(uid) == 0, L.rs_nccl_last_error()
    comm = C.c_void_p()
    assert L.rs_comm_init_rank(1, 0, uid, 0, C.byref(comm)) == 0, L.rs_nccl_last_error()
    red = np.zeros((4, 5), dtype=np.uint64)
    assert L.rs_reduce_stats(g._h, comm, 0, red.ctypes.data_as(C.c_void_p)) == 0, L.rs_nccl_last_error()
    L.rs_comm_destroy(comm)
    assert np.array_equal(red, g.get_stats()) and red[0].sum() > 0
    g.close()


def test_rs_batch_two_gpus_equal_one(tmp_path):
    """rs_batch --gpus 2 (one host thread per GPU, ncclCommInitAll, one ncclReduce at the end): the totals equal the
    single-GPU run of the same cells."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    exe = os.path.join(ROOT, "radiosaber_b200", "rs_batch")
    cfg = os.path.join(ROOT, "tests", "data", "cfg20x5.json")
    runs = []
    for n in (1, 2):
        r = subprocess.run([exe, "--algo", "9", "--config", cfg, "--cells", "301", "--ttis", "64", "--seed", "3",
                            "--gpus", str(n), "--log-cell", "200", "--log-prefix", str(tmp_path / f"c{n}")],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-600:]
        runs.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert runs[0]["slice_bytes"] == runs[1]["slice_bytes"] and runs[0]["slice_rbs"] == runs[1]["slice_rbs"]
    assert sum(runs[0]["slice_bytes"]) > 0 and runs[1]["gpus"] == 2
    assert (tmp_path / "c1.stderr").read_text() == (tmp_path / "c2.stderr").read_text()   # cell 200 lives on GPU 1
    assert (tmp_path / "c1.stdout").read_text() == (tmp_path / "c2.stdout").read_text()


def test_two_half_batches_equal_one_launch():
    """A big batch runs as two half-batches on two streams (each half's launches chained, the tail of one launch
    overlapping the head of the next); RS_NO_SPLIT=1 keeps one launch per step.  Same results, both through the
    host-buffer path and with everything resident on the device."""
    import torch
    S, n, B, T = 20, 5, 2400, 11
    w, p, u2s = _cfg(S, n, mix=True)
    U, G = len(u2s), 64
    cqi = workload.synth_cqi(21, 0, B, 0, T, U, G)
    r2 = workload.synth_rand2(21, 0, B, 0, T, S)
    _, dts = workload.tti_clock(T)
    res = []
    for no_split in (False, True):
        if no_split:
            os.environ["RS_NO_SPLIT"] = "1"
        try:
            g = sched.Scheduler(9, w, p, u2s, B)
        finally:
            os.environ.pop("RS_NO_SPLIT", None)
        a = g.run_host(cqi, r2, dts, want_aux=True, ttis_per_launch=4)
        st_host = g.get_state()
        g.reset_state()
        dev = torch.device("cuda", 0)
        d_cqi = torch.from_numpy(cqi).to(dev)
        d_r2 = torch.from_numpy(r2).to(dev)
        d_rbg = torch.empty((T, B, G), dtype=torch.int16, device=dev)
        d_bits = torch.empty((T, B, U), dtype=torch.int32, device=dev)
        g.run_device(T, d_cqi.data_ptr(), B * U * G, d_r2.data_ptr(), dts, {"rbg_to_ue": d_rbg.data_ptr(), "tbs_bits": d_bits.data_ptr()},
                     ttis_per_launch=4)
        st_dev = g.get_state()            # synchronises the handle's stream: must be behind BOTH halves
        res.append((a, st_host, d_rbg.cpu().numpy(), d_bits.cpu().numpy(), st_dev))
        g.close()
    (a0, s0, r0, b0, t0), (a1, s1, r1, b1, t1) = res
    for k in a0:
        assert np.array_equal(a0[k], a1[k]), k
    for k in s0:
        assert np.array_equal(s0[k], s1[k]), k
        assert np.array_equal(t0[k], t1[k]), k
        assert np.array_equal(s0[k], t0[k]), k         # host-buffer path == device-resident path
    assert np.array_equal(r0, r1) and np.array_equal(b0, b1)
    assert np.array_equal(r0, a0["rbg_to_ue"]) and np.array_equal(b0, a0["tbs_bits"])
