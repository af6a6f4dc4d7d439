"""Replay of the long-horizon golden records (tests/golden/long/*.npz, tools/make_golden_long.py): the inputs are
regenerated from the record's seed (synthetic CQI) or unpacked from its CQI slabs (trace-driven), every block of
`block` TTIs is hashed like the generator hashed the reference's output, and the state at each block's end is
compared with the record's checkpoint.  Test infrastructure, shared by the CPU (oracle) and GPU (CUDA) tests."""
from __future__ import annotations

import hashlib
import os

import numpy as np

from radiosaber_b200 import workload
from tools import golden_io

LONG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "long")


def long_names():
    return sorted(f[:-4] for f in os.listdir(LONG_DIR) if f.endswith(".npz")) if os.path.isdir(LONG_DIR) else []


def load_long(name):
    return golden_io.load_npz(os.path.join(LONG_DIR, name + ".npz"))


def cqi_of_tti(rec, t):
    """uint8 [U][G] the reference's scheduler saw at TTI t."""
    U, G = int(rec["U"]), int(rec["G"])
    if rec["source"] == "synth":
        return workload.synth_cqi(int(rec["seed"]), 0, 1, t, 1, U, G)[0, 0]
    k = int(np.searchsorted(rec["slab_start"], t, side="right")) - 1
    p = rec["slabs"][k]
    out = np.empty((U, G), dtype=np.uint8)
    out[:, 0::2] = p & 15
    out[:, 1::2] = p >> 4
    return out


def replay_long(sched, rec, blocks=None):
    """Drive a 1-cell scheduler (oracle or CUDA: same python surface) through the record.  blocks: iterable of block
    indices to replay, each started from the previous block's checkpoint (None = the whole run, chained on the
    scheduler's own state).  Returns a list of (block, what) mismatches."""
    algo, T, blk = int(rec["algo"]), int(rec["T"]), int(rec["block"])
    S = int(rec["S"])
    rand2 = workload.synth_rand2(int(rec["seed"]), 0, 1, 0, T, S)[:, 0, :]
    transport, nvs = algo in (8, 9, 10, 101, 103), algo in (7, 11)
    bad = []

    def restore(b):
        if b == 0:
            avg, tx, st = rec["avg_before0"], rec["tx_before0"], rec["state_before0"]
            cb = cr = np.zeros(int(rec["U"]), dtype=np.uint64)
        else:
            avg, tx, st = rec["ck_avg_after"][b - 1], rec["ck_tx_after"][b - 1], rec["ck_state_after"][b - 1]
            cb, cr = rec["ck_cum_bytes"][b - 1], rec["ck_cum_rbs"][b - 1]
        sched.set_state(avg_rate=avg[None], tx_bytes=tx[None], cum_bytes=cb[None], cum_rbs=cr[None],
                        slice_offset=st[None] if transport else None, nvs_ewma=st[None] if nvs else None)

    todo = range(T // blk) if blocks is None else blocks
    prev = None
    for b in todo:
        if blocks is not None or b == 0:
            restore(b)
        elif prev is not None and prev != b - 1:
            restore(b)
        h = hashlib.sha256()
        st = None
        for t in range(b * blk, (b + 1) * blk):
            out = sched.step(cqi_of_tti(rec, t)[None], rand2[t][None], dt=float(rec["dt"][t]))
            st = sched.get_state()
            h.update(np.ascontiguousarray(out["rbg_to_ue"][0], dtype=np.int16).tobytes())
            h.update(np.ascontiguousarray(out["tbs_bits"][0], dtype=np.int32).tobytes())
            h.update(np.ascontiguousarray(st["avg_rate"][0], dtype=np.float64).tobytes())
        if not np.array_equal(np.frombuffer(h.digest(), dtype=np.uint8), rec["sha256"][b]):
            bad.append((b, "sha256(rbg_to_ue|bits|avg_rate)"))
        for k, ck in (("avg_rate", "ck_avg_after"), ("tx_bytes", "ck_tx_after"), ("cum_bytes", "ck_cum_bytes"),
                      ("cum_rbs", "ck_cum_rbs")):
            if not np.array_equal(st[k][0], rec[ck][b]):
                bad.append((b, k))
        if transport and not np.array_equal(st["slice_offset"][0], rec["ck_state_after"][b]):
            bad.append((b, "slice_offset"))
        if nvs and not np.array_equal(st["nvs_ewma"][0], rec["ck_state_after"][b]):
            bad.append((b, "nvs_ewma"))
        prev = b
    return bad
