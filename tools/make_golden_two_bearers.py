#!/usr/bin/env python3
"""tests/golden/two_bearers/a<id>.npz: per-BEARER records of the UNMODIFIED reference on a cell whose internet-flow
slices carry two bearers per UE (tests/data/cfg_two_bearers.json; MAX_BEARERS = 2, packet-scheduler.h:31), for the batch
(device-state) path: every bearer's dataToTransmit and head-of-line delay per TTI (ref_harness --bearer-log), the
allocation the reference made, and every bearer's EWMA rate / byte and RB counters after each TTI.  CQI and rand()
inputs are regenerated from the seed.  Needs oracle/_ref/ref_harness (i.e. /root/reference mounted)."""
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
from tools.make_golden_logs_two_bearers import CFG, HARNESS, n_bearers  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "two_bearers")
TTIS, SEED = 200, 43
IDS = (9, 8, 7, 10, 101, 103, 11, 1)


def main():
    os.makedirs(OUT, exist_ok=True)
    cfg = json.load(open(CFG))
    U, S = sum(cfg["ues_per_slice"]), len(cfg["ues_per_slice"])
    nbr = n_bearers(cfg)
    with tempfile.TemporaryDirectory() as tmp:
        cqi, rnd = os.path.join(tmp, "cqi.bin"), os.path.join(tmp, "rand.bin")
        workload.synth_cqi(SEED, 0, 1, 0, TTIS, U, 64)[:, 0].tofile(cqi)
        workload.synth_rand2(SEED, 0, 1, 0, TTIS, S)[:, 0, :].astype("<i4").tofile(rnd)
        for algo in IDS:
            rec_path, bl, al = os.path.join(tmp, f"a{algo}.bin"), os.path.join(tmp, f"a{algo}.blog"), os.path.join(tmp, f"a{algo}.alog")
            cmd = [HARNESS, "--algo", str(algo), "--config", CFG, "--ttis", str(TTIS), "--seed", str(SEED), "--cqi", cqi,
                   "--rand", rnd, "--bearers", str(nbr), "--out", rec_path, "--bearer-log", bl]
            if algo == 10:
                cmd += ["--alloc-log", al]
            rl = os.path.join(tmp, f"a{algo}.rlog")
            if algo == 11:
                cmd += ["--rand-log", rl]
            subprocess.run(cmd, check=True, capture_output=True)
            rec = golden_io.parse_record_stream(rec_path)
            T = rec["T"]
            assert T == TTIS
            raw = open(bl, "rb").read()
            pre = np.dtype([("user", "<i4"), ("prio", "<i4"), ("data", "<i4"), ("tx", "<i4"), ("hol", "<f8"), ("avg", "<f8")])
            post = np.dtype([("avg", "<f8"), ("tx", "<i4"), ("pad", "<i4"), ("cum_bytes", "<u8"), ("cum_rbs", "<u8")])
            out = {k: np.zeros((T, U, 2), dt) for k, dt in (("queue", np.int32), ("hol", np.float64), ("avg_before", np.float64),
                                                              ("tx_before", np.int32), ("avg_after", np.float64), ("tx_after", np.int32),
                                                              ("cum_bytes", np.uint64), ("cum_rbs", np.uint64))}
            exists = np.zeros((U, 2), np.uint8)
            app_id = np.full((U, 2), -1, np.int32)
            pos = 0
            for t in range(T):
                n = int(np.frombuffer(raw, "<i4", 1, pos)[0])
                pos += 4
                assert n == nbr
                a = np.frombuffer(raw, pre, n, pos)
                pos += pre.itemsize * n
                b = np.frombuffer(raw, post, n, pos)
                pos += post.itemsize * n
                u, i = a["user"], a["prio"]
                # container order inside a UE is priority order (the scenario starts flow j with priority j in order)
                assert all(np.all(np.diff(i[u == x]) > 0) for x in np.unique(u))
                exists[u, i] = 1
                app_id[u, i] = np.arange(n)     # applications are created UE by UE, flow by flow: id = container index
                out["queue"][t, u, i], out["hol"][t, u, i] = a["data"], a["hol"]
                out["avg_before"][t, u, i], out["tx_before"][t, u, i] = a["avg"], a["tx"]
                out["avg_after"][t, u, i], out["tx_after"][t, u, i] = b["avg"], b["tx"]
                out["cum_bytes"][t, u, i], out["cum_rbs"][t, u, i] = b["cum_bytes"], b["cum_rbs"]
            assert pos == len(raw)
            dt = rec["now"][:, None] - rec["last_update"]
            assert (dt == dt[:, :1]).all()
            g = {"algo": algo, "S": S, "U": U, "G": rec["G"], "T": T, "seed": SEED, "config_json": json.dumps(cfg),
                 "weight": rec["weight"], "params": rec["params"], "ue_to_slice": rec["ue_to_slice"], "dt": dt[:, 0].copy(),
                 "exists": exists, "app_id": app_id, "rbg_to_ue": rec["rbg_to_ue"], "bits": rec["bits"],
                 "final_cqi": rec["final_cqi"], "target": rec["target"], "quota": rec["quota"], "nvs_slice": rec["nvs_slice"],
                 "state_before": rec["state_before"], "state_after": rec["state_after"], "rand2": rec["rand2"]}
            g.update(out)
            if algo == 10:
                rawa = np.fromfile(al, dtype="<i2")
                G = int(rec["G"])
                ue = np.full((T, 2 * G), -1, np.int16)
                rb = np.full((T, 2 * G), -1, np.int16)
                cnt = np.zeros(T, np.int32)
                p2 = 0
                for t in range(T):
                    k = int(rawa[p2]) | (int(rawa[p2 + 1]) << 16)
                    p2 += 2
                    pairs = rawa[p2:p2 + 2 * k].reshape(k, 2)
                    p2 += 2 * k
                    cnt[t] = k
                    ue[t, :k], rb[t, :k] = pairs[:, 0], pairs[:, 1]
                g["alloc_n"], g["alloc_ue"], g["alloc_rbg"] = cnt, ue, rb
            if algo == 11:   # every rand() value of the 300-sample search, TTI by TTI (downlink-nvs-scheduler.cpp:437-446)
                rawr = np.fromfile(rl, dtype="<i4")
                draws, p3 = [], 0
                for _ in range(T):
                    k = int(rawr[p3])
                    draws.append(rawr[p3 + 1:p3 + 1 + k])
                    p3 += 1 + k
                assert p3 == len(rawr)
                width = max(len(x) for x in draws)
                g["rand_ng_n"] = np.array([len(x) for x in draws], dtype=np.int32)
                g["rand_ng"] = np.stack([np.pad(x, (0, width - len(x))) for x in draws]).astype(np.int32)
            both = int(((out["queue"][:, :, 0] > 0) & (out["queue"][:, :, 1] > 0)).sum())
            handover = int(((out["queue"][:, :, 1] == 0) & (out["queue"][:, :, 0] > 0) & exists[None, :, 1].astype(bool)).sum())
            path = os.path.join(OUT, f"a{algo}.npz")
            golden_io.save_npz(path, g)
            print(f"id {algo}: {T} TTIs, {both} (TTI, UE) with both bearers queued, {handover} with only the low-priority one "
                  f"-> {os.path.getsize(path) // 1024} KiB", flush=True)


if __name__ == "__main__":
    main()
