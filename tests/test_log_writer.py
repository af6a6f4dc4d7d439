"""Log-compatible writer: the text the reference's schedulers print per TTI, regenerated
from batch results, against the unmodified reference's own output (tests/golden/logs/, captured by
oracle/ref_harness.cpp --log-out through tools/make_golden_logs.py).  Host-only code: runs without a GPU;
the GPU test feeds it with what the CUDA path returns."""
import os

import numpy as np
import pytest

from radiosaber_b200 import sched
from tests.helpers import GOLDEN, load_golden

LOGS = os.path.join(GOLDEN, "logs")
CASES = sorted(f[:-7] for f in os.listdir(LOGS) if f.endswith(".stdout"))
FIRST_TS = 100   # PacketScheduler::m_ts at the first TTI with bearers (applications start at 0.1 s)


def _ref_text(name):
    """The reference's text of the recorded TTIs.  With finite flows its stderr also carries the RLC's own
    "ipflow end app: ... fct: ..." lines (um-rlc-entity.cpp:154-159, printed from TransmissionProcedure): not
    scheduler output, dropped here."""
    out = open(os.path.join(LOGS, name + ".stdout")).read()
    err = open(os.path.join(LOGS, name + ".stderr")).read()
    err = "".join(l for l in err.splitlines(keepends=True) if not l.startswith("ipflow "))
    return out, err


def _n_ttis(name, rec):
    out, err = _ref_text(name)
    if int(rec["algo"]) == 1:
        return len({l.split()[0] for l in err.splitlines() if l[:1].isdigit()})
    return sum(1 for l in out.splitlines() if l.strip().isdigit())


@pytest.mark.parametrize("name", CASES)
def test_writer_reproduces_reference_text_from_reference_results(name):
    rec = load_golden(name)
    algo = int(rec["algo"])
    T = _n_ttis(name, rec)
    assert T >= 4
    lw = sched.LogWriter(algo, rec["ue_to_slice"], int(rec["S"]), cqi_per_rb=int(rec["cqi_per_rb"]))
    tq = algo in (8, 9, 101, 103)
    for t in range(T):
        kw = {"queue": rec["queue"][t], "hol": rec["hol"][t]} if "queue" in rec else {}
        if algo == 10:   # the grant list; the record holds it user by user, the order of the RB lists
            lw.tti_grants(FIRST_TS + t, rec["cqi"][t], int(rec["alloc_n"][t]), rec["alloc_ue"][t], rec["alloc_rbg"][t],
                          rec["bits"][t], rec["final_cqi"][t], rec["target"][t], rec["quota"][t], **kw)
            continue
        lw.tti(FIRST_TS + t, rec["cqi"][t], rec["rbg_to_ue"][t], rec["bits"][t], rec["final_cqi"][t],
               rec["target"][t] if tq else None, rec["quota"][t] if tq else None, **kw)
    out, err = _ref_text(name)
    assert lw.stdout == out
    assert lw.stderr == err
    cb, cr = lw.counters()
    assert np.array_equal(cb, rec["cum_bytes"][T - 1]) and np.array_equal(cr, rec["cum_rbs"][T - 1])
    lw.clear()
    assert lw.stdout == "" and lw.stderr == ""


@pytest.mark.parametrize("algo", [9, 8, 7, 10, 101, 103, 11])
def test_writer_two_bearers_per_ue_batch_mode(algo):
    """A cell whose internet-flow slices carry two bearers per UE (MAX_BEARERS = 2): per-bearer queues, head-of-line
    delays and application ids in, the unmodified reference's text of the same 200-TTI run out
    (tools/make_golden_logs_two_bearers_batch.py; the per-bearer record is tests/golden/two_bearers/a<id>.npz)."""
    import gzip
    from radiosaber_b200 import workload
    d = os.path.join(GOLDEN, "two_bearers")
    rec = np.load(os.path.join(d, f"a{algo}.npz"))
    T, U, S, G = int(rec["T"]), int(rec["U"]), int(rec["S"]), int(rec["G"])
    cqi = workload.synth_cqi(int(rec["seed"]), 0, 1, 0, T, U, G)[:, 0]
    lw = sched.LogWriter(algo, rec["ue_to_slice"], S, n_bearers=2, app_ids=rec["app_id"])
    tq = algo in (8, 9, 101, 103)
    for t in range(T):
        kw = {"queue": rec["queue"][t], "hol": rec["hol"][t]}
        if algo == 10:
            lw.tti_grants(FIRST_TS + t, cqi[t], int(rec["alloc_n"][t]), rec["alloc_ue"][t], rec["alloc_rbg"][t],
                          rec["bits"][t], rec["final_cqi"][t], rec["target"][t], rec["quota"][t], **kw)
        else:
            lw.tti(FIRST_TS + t, cqi[t], rec["rbg_to_ue"][t], rec["bits"][t], rec["final_cqi"][t],
                   rec["target"][t] if tq else None, rec["quota"][t] if tq else None, **kw)
    out = gzip.open(os.path.join(d, f"a{algo}.stdout.gz"), "rt").read()
    err = gzip.open(os.path.join(d, f"a{algo}.stderr.gz"), "rt").read()
    err = "".join(l for l in err.splitlines(keepends=True) if not l.startswith("ipflow "))
    assert lw.stdout == out
    assert lw.stderr == err
    cb, cr = lw.counters()
    assert np.array_equal(cb, rec["cum_bytes"][T - 1]) and np.array_equal(cr, rec["cum_rbs"][T - 1])
    # some TTI served both bearers of a UE, and some bearer was served while the other one had nothing queued
    both = {}
    for l in err.splitlines():
        f = l.split()
        if len(f) > 10 and f[1] == "app:":
            both[(f[0], f[10])] = both.get((f[0], f[10]), 0) + 1
    assert max(both.values()) == 2


def test_writer_rejects_two_bearers_for_flow_level_pf():
    with pytest.raises(Exception):
        sched.LogWriter(1, np.zeros(4, np.int32), 1, n_bearers=2)


def test_writer_layouts_agree():
    rec = load_golden("a9_fix20x5_synth")
    texts = []
    for layout in (0, 1, 2):
        lw = sched.LogWriter(9, rec["ue_to_slice"], int(rec["S"]), cqi_per_rb=layout)
        cqi = rec["cqi"][0]
        c = {0: cqi, 1: np.repeat(cqi, 8, axis=1), 2: sched.pack_cqi(cqi)}[layout]
        lw.tti(100, c, rec["rbg_to_ue"][0], rec["bits"][0], rec["final_cqi"][0], rec["target"][0], rec["quota"][0])
        texts.append((lw.stdout, lw.stderr))
    assert texts[0] == texts[1] == texts[2]


def test_plotter_style_parse_of_stderr():
    """The fields plot_throughput.py:35-47 reads (ts, app, cumu_bytes, cumu_rbs, slice) round-trip."""
    rec = load_golden("a8_fix20x5_synth")
    lw = sched.LogWriter(8, rec["ue_to_slice"], int(rec["S"]))
    for t in range(4):
        lw.tti(100 + t, rec["cqi"][t], rec["rbg_to_ue"][t], rec["bits"][t], rec["final_cqi"][t], rec["target"][t],
               rec["quota"][t])
    last = {}
    for line in lw.stderr.splitlines():
        w = line.split()
        if len(w) > 2 and w[1] == "app:":
            last[int(w[2])] = (int(w[4]), int(w[6]), int(w[12]))
    for u, (b, r, s) in last.items():
        assert b == int(rec["cum_bytes"][3][u]) and r == int(rec["cum_rbs"][3][u]) and s == int(rec["ue_to_slice"][u])


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_cuda_results_reproduce_reference_text(name):
    """End to end: CUDA path -> results -> writer == the reference's own log text."""
    rec = load_golden(name)
    algo = int(rec["algo"])
    T = _n_ttis(name, rec)
    g = sched.Scheduler(algo, rec["weight"], rec["params"], rec["ue_to_slice"], 1, cqi_per_rb=int(rec["cqi_per_rb"]))
    tq = algo in (8, 9, 10, 101, 103)
    g.set_state(avg_rate=rec["avg_before"][0][None], tx_bytes=rec["tx_before"][0][None],
                slice_offset=rec["state_before"][0][None] if tq else None,
                nvs_ewma=rec["state_before"][0][None] if algo in (7, 11) else None)
    qkw = {"queue": rec["queue"][:T, None], "hol": rec["hol"][:T, None]} if "queue" in rec else {}
    draws = rec["rand_ng"] if algo == 11 else rec["rand2"]
    res = g.run_host(rec["cqi"][:T, None], draws[:T, None, :], rec["dt"][:T], want_aux=True, **qkw)
    lw = sched.LogWriter(algo, rec["ue_to_slice"], int(rec["S"]), cqi_per_rb=int(rec["cqi_per_rb"]))
    for t in range(T):
        kw = {"queue": rec["queue"][t], "hol": rec["hol"][t]} if "queue" in rec else {}
        if algo == 10:
            lw.tti_grants(FIRST_TS + t, rec["cqi"][t], int(res["alloc_n"][t, 0]), res["alloc_ue"][t, 0], res["alloc_rbg"][t, 0],
                          res["tbs_bits"][t, 0], res["final_cqi"][t, 0], res["slice_target"][t, 0], res["slice_quota"][t, 0], **kw)
            continue
        lw.tti(FIRST_TS + t, rec["cqi"][t], res["rbg_to_ue"][t, 0], res["tbs_bits"][t, 0], res["final_cqi"][t, 0],
               res["slice_target"][t, 0] if tq else None,
               res["slice_quota"][t, 0] if tq else None, **kw)
    out, err = _ref_text(name)
    assert lw.stdout == out and lw.stderr == err
    st = g.get_state()
    cb, cr = lw.counters()
    assert np.array_equal(cb, st["cum_bytes"][0]) and np.array_equal(cr, st["cum_rbs"][0])
    g.close()


def test_grant_list_call_belongs_to_id_10():
    """rs_log_tti wants the single-valued RBG->UE map (every id but 10), rs_log_tti_grants the grant list (id 10 only)."""
    rec = load_golden("a10_fix20x5_synth")
    U, S, G = int(rec["U"]), int(rec["S"]), int(rec["G"])
    lw10 = sched.LogWriter(10, rec["ue_to_slice"], S)
    with pytest.raises(sched.RsError, match="rs_log_tti_grants"):
        lw10.tti(100, rec["cqi"][0], rec["rbg_to_ue"][0], rec["bits"][0], rec["final_cqi"][0], rec["target"][0], rec["quota"][0])
    lw9 = sched.LogWriter(9, rec["ue_to_slice"], S)
    with pytest.raises(sched.RsError, match="id 10"):
        lw9.tti_grants(100, rec["cqi"][0], int(rec["alloc_n"][0]), rec["alloc_ue"][0], rec["alloc_rbg"][0], rec["bits"][0],
                       rec["final_cqi"][0], rec["target"][0], rec["quota"][0])
    bad = rec["alloc_ue"][0].copy()
    bad[0] = U   # a user outside the cell
    with pytest.raises(sched.RsError, match="grant 0"):
        lw10.tti_grants(100, rec["cqi"][0], int(rec["alloc_n"][0]), bad, rec["alloc_rbg"][0], rec["bits"][0],
                        rec["final_cqi"][0], rec["target"][0], rec["quota"][0])
    assert lw10.stdout == "" and lw10.stderr == ""
