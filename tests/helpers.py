"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import os

import numpy as np

from tools import golden_io

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz"))


def load_golden(name):
    return golden_io.load_npz(os.path.join(GOLDEN, name + ".npz"))


def grants_by_user(ue, rbg):
    """{user: [RBGs in grant order]} from parallel grant arrays."""
    d = {}
    for u, g in zip(np.asarray(ue).tolist(), np.asarray(rbg).tolist()):
        d.setdefault(u, []).append(g)
    return d


def replay_golden(sched, rec, check_final_cqi=True):
    """Drive a 1-cell scheduler (oracle or CUDA, same python surface) through a golden record,
    chaining its own state, and compare every TTI with what the reference produced.
    Returns the list of mismatching (tti, field) pairs."""
    algo, T = int(rec["algo"]), int(rec["T"])
    bad = []
    sched.set_state(avg_rate=rec["avg_before"][0][None], tx_bytes=rec["tx_before"][0][None],
                    slice_offset=rec["state_before"][0][None] if algo in (8, 9, 10, 101, 103) else None,
                    nvs_ewma=rec["state_before"][0][None] if algo in (7, 11) else None)
    for t in range(T):
        # id 11: every rand() value of the 300-sample search (downlink-nvs-scheduler.cpp:437-446)
        draws = rec["rand_ng"][t][None] if algo == 11 else rec["rand2"][t][None]
        kw = {}
        if "queue" in rec:   # queue-aware records: every bearer's dataToTransmit and head-of-line delay
            kw = {"queue": rec["queue"][t][None], "hol": rec["hol"][t][None]}
        out = sched.step(rec["cqi"][t][None], draws, dt=float(rec["dt"][t]), want_aux=True, **kw)
        st = sched.get_state()

        def chk(field, a, b):
            if not np.array_equal(np.asarray(a), np.asarray(b)):
                bad.append((t, field))

        chk("rbg_to_ue", out["rbg_to_ue"][0], rec["rbg_to_ue"][t])
        if algo == 10:   # every (user, RBG) grant, per user in the order of the user's RB list
            n, m = int(out["alloc_n"][0]), int(rec["alloc_n"][t])
            got = grants_by_user(out["alloc_ue"][0][:n], out["alloc_rbg"][0][:n])
            want = grants_by_user(rec["alloc_ue"][t][:m], rec["alloc_rbg"][t][:m])
            if n != m or got != want:
                bad.append((t, "grants"))
        chk("tbs_bits", out["tbs_bits"][0], rec["bits"][t])
        if check_final_cqi and algo != 1 and "final_cqi" in out:  # PF prints no final_cqi line
            chk("final_cqi", out["final_cqi"][0], rec["final_cqi"][t])
        chk("avg_rate", st["avg_rate"][0], rec["avg_after"][t])
        chk("tx_bytes", st["tx_bytes"][0], rec["tx_after"][t])
        chk("cum_bytes", st["cum_bytes"][0], rec["cum_bytes"][t])
        chk("cum_rbs", st["cum_rbs"][0], rec["cum_rbs"][t])
        if algo in (8, 9, 10, 101, 103):
            if "slice_target" in out:
                chk("slice_target", out["slice_target"][0], rec["target"][t])
                chk("slice_quota", out["slice_quota"][0], rec["quota"][t])
            chk("slice_offset", st["slice_offset"][0], rec["state_after"][t])
        if algo in (7, 11):
            chk("nvs_ewma", st["nvs_ewma"][0], rec["state_after"][t])
            if "nvs_slice" in out:
                chk("nvs_slice", out["nvs_slice"][0], rec["nvs_slice"][t])
    return bad
