#!/usr/bin/env python3
"""Long-horizon golden records from the UNMODIFIED reference (oracle/_ref/ref_harness): thousands of TTIs of the
shapes bench.py times, small enough to commit.

A record keeps the inputs as a seed (synthetic CQI, regenerated with radiosaber_b200.workload) or as the distinct
CQI slabs the simulator ingested from cqi-traces-noise0 (one per CQI report, 4-bit packed), the initial state, and
per block of BLOCK TTIs: the SHA-256 of what the reference produced (rbg_to_ue | allocated bits | EWMA rates, TTI by
TTI) and a checkpoint of the state at the block's end, so that a test can replay the whole run or any block of it.

Needs /root/reference and `make -C oracle ref`.  Usage: python tools/make_golden_long.py [--only NAME]
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosaber_b200 import workload  # noqa: E402
from tools import golden_io  # noqa: E402

HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
REF_EXP = "/root/reference/NSDI23-radiosaber-experiments"
OUT_DIR = os.path.join(ROOT, "tests", "golden", "long")
BLOCK = 100
FIX20X5 = f"{REF_EXP}/exp-fix20slices/5ues/config-pf.json"
DIFFW = f"{REF_EXP}/exp-customization/exp-backlogged-20slicesdiffw/config.json"

# name, algo, config, TTIs, CQI source, seed
CASES = [
    # the headline shape of bench.py (20 slices x 5 UEs, PF, fresh synthetic CQI every TTI), past the TTIs it times
    ("a9_fix20x5_synth_long", 9, FIX20X5, 1300, "synth", 2),
    ("a8_fix20x5_synth_long", 8, FIX20X5, 1000, "synth", 3),
    ("a7_fix20x5_synth_long", 7, FIX20X5, 1000, "synth", 4),
    ("a1_fix20x5_synth_long", 1, FIX20X5, 1000, "synth", 5),
    # BASELINE.json configs[0] at full length: SingleCellWithI 1 9 1 30 1 12 <diffw config>, mapping1.config
    ("a9_diffw_trace_12s", 9, DIFFW, 12000, "trace", 1),
]


def stream_records(path):
    """Yield (header, per-TTI dict) from a ref_harness --out file without loading it whole."""
    f = open(path, "rb")
    if f.read(8) != golden_io.MAGIC:
        raise ValueError("bad magic")
    algo, S, U, R, rbg = (int(x) for x in np.frombuffer(f.read(20), "<i4"))
    G = R // rbg
    hdr = {"algo": algo, "S": S, "U": U, "R": R, "rbg_size": rbg, "G": G,
           "weight": np.frombuffer(f.read(8 * S), "<f8").copy(),
           "params": np.frombuffer(f.read(16 * S), "<i4").reshape(S, 4).copy(),
           "ue_to_slice": np.frombuffer(f.read(4 * U), "<i4").copy()}
    fields = [
        ("now", "<f8", 1), ("avg_before", "<f8", U), ("tx_before", "<i4", U), ("last_update", "<f8", U),
        ("state_before", "<f8", S), ("cqi_rb", "u1", U * R), ("active", "u1", U), ("rand2", "<i4", 2),
        ("rbg_to_ue", "<i2", G), ("bits", "<i4", U), ("final_cqi", "u1", U),
        ("target", "<i4", S), ("quota", "<i4", S), ("nvs_slice", "<i4", 1),
        ("avg_after", "<f8", U), ("tx_after", "<i4", U), ("cum_bytes", "<u8", U), ("cum_rbs", "<u8", U),
        ("state_after", "<f8", S),
    ]
    yield hdr
    t = 0
    while True:
        head = f.read(8)
        if not head:
            return
        mark, tti = (int(x) for x in np.frombuffer(head, "<i4"))
        if mark != golden_io.TTI_MARK or tti != t:
            raise ValueError("bad TTI header")
        rec = {}
        for name, dt, n in fields:
            rec[name] = np.frombuffer(f.read(np.dtype(dt).itemsize * n), dt)
        yield rec
        t += 1


def block_digest(parts):
    h = hashlib.sha256()
    for p in parts:
        h.update(np.ascontiguousarray(p).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8)


def run_case(name, algo, config, n_ttis, source, seed, tmp):
    cfg = json.load(open(config))
    U = int(sum(cfg["ues_per_slice"]))
    S = len(cfg["ues_per_slice"])
    rec_path = os.path.join(tmp, name + ".bin")
    rand2 = workload.synth_rand2(seed, 0, 1, 0, n_ttis, S)[:, 0, :]
    rand_path = os.path.join(tmp, name + ".rand")
    rand2.astype("<i4").tofile(rand_path)
    cmd = [HARNESS, "--algo", str(algo), "--config", config, "--ttis", str(n_ttis), "--out", rec_path,
           "--seed", str(seed), "--rand", rand_path]
    if source == "synth":
        cqi_path = os.path.join(tmp, name + ".cqi")
        with open(cqi_path, "wb") as f:
            for t0 in range(0, n_ttis, 100):
                workload.synth_cqi(seed, 0, 1, t0, min(100, n_ttis - t0), U, 64)[:, 0].tofile(f)
        cmd += ["--cqi", cqi_path]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    it = stream_records(rec_path)
    hdr = next(it)
    G, rbg = hdr["G"], hdr["rbg_size"]
    dts, nows, digests, ck = [], [], [], {k: [] for k in ("avg_after", "tx_after", "cum_bytes", "cum_rbs", "state_after")}
    slabs, slab_start, parts, last = [], [], [], None
    first = None
    T = 0
    for r in it:
        if first is None:
            first = {k: r[k].copy() for k in ("avg_before", "tx_before", "state_before")}
        dt = float(r["now"][0]) - r["last_update"]
        assert (dt == dt[0]).all(), "bearers disagree on lastUpdate"
        assert r["active"].all(), "a bearer without packets in a backlogged run"
        if algo in (8, 9):
            assert (r["rand2"] == rand2[T]).all(), "scripted rand() values were not the ones consumed"
        dts.append(dt[0])
        nows.append(float(r["now"][0]))
        c = r["cqi_rb"].reshape(hdr["U"], G, rbg)
        assert (c == c[..., :1]).all(), "CQI varies inside an RBG"
        c = c[..., 0]
        if source == "trace" and (last is None or not np.array_equal(c, last)):
            slabs.append((c[:, 0::2] | (c[:, 1::2] << 4)).astype(np.uint8))
            slab_start.append(T)
            last = c.copy()
        if source == "synth":
            assert np.array_equal(c, workload.synth_cqi(seed, 0, 1, T, 1, U, 64)[0, 0]), "injected CQI not consumed"
        parts += [r["rbg_to_ue"], r["bits"], r["avg_after"]]
        T += 1
        if T % BLOCK == 0:
            digests.append(block_digest(parts))
            parts = []
            for k in ck:
                ck[k].append(r[k].copy())
    assert T == n_ttis and not parts, (T, n_ttis)
    out = dict(hdr)
    out.update({"T": T, "block": BLOCK, "source": source, "seed": seed, "config_json": json.dumps(cfg),
                "dt": np.array(dts), "now": np.array(nows), "sha256": np.stack(digests),
                "avg_before0": first["avg_before"], "tx_before0": first["tx_before"], "state_before0": first["state_before"]})
    for k in ck:
        out["ck_" + k] = np.stack(ck[k])
    if source == "trace":
        out["slab_start"] = np.array(slab_start, dtype=np.int32)
        out["slabs"] = np.stack(slabs)
    os.makedirs(OUT_DIR, exist_ok=True)
    path = os.path.join(OUT_DIR, name + ".npz")
    golden_io.save_npz(path, out)
    print(f"{name}: {T} TTIs, {len(digests)} blocks, {len(slabs)} CQI slabs -> {os.path.getsize(path) / 1024:.0f} KiB", flush=True)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--tmp", default=None, help="scratch directory for the raw record streams (1.4 GB for the 12-s run)")
    args = ap.parse_args()
    if not os.path.exists(HARNESS):
        print("oracle/_ref/ref_harness missing: run `make -C oracle ref` first", file=sys.stderr)
        return 1
    with tempfile.TemporaryDirectory(dir=args.tmp) as tmp:
        for case in CASES:
            if args.only and case[0] != args.only:
                continue
            run_case(*case, tmp)
    return 0


if __name__ == "__main__":
    sys.exit(main())
