#!/usr/bin/env python3
"""Generate tests/golden/*.npz by running the UNMODIFIED reference.

Needs oracle/_ref/ref_harness (``make -C oracle ref``; only possible where
/root/reference is mounted, i.e. in the build container).  Each case runs the
reference's own SingleCellWithI scenario with one of its downlink schedulers
and records, per TTI, everything the scheduler consumed (CQI, EWMA rates,
slice offsets / NVS credits, the two rand() draws, Now - lastUpdate) and
produced (RBG->UE map, allocated bits, final CQI, targets/quotas, updated
state, cumulative bytes/RBs).  The committed fixtures make the parity tests
runnable where the reference is absent (the GPU box).

Usage: python tools/make_golden.py [--only NAME]
"""
from __future__ import annotations

import argparse
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
GOLDEN = os.path.join(ROOT, "tests", "golden")

SMALL_CFG = {
    # 5 slices, few UEs: single UEs end up with 15/20/25 RBGs, which exercises the
    # TransportBlockSizeTable[-1] read of AMCModule.cpp:312-316
    "slices": [
        {"n_slices": 2, "weight": 0.3, "video_app": 0, "video_bitrate": 0, "internet_flow": 0, "if_bitrate": 0,
         "backlog_flow": 1, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 1},
        {"n_slices": 3, "weight": 0.1333333333333, "video_app": 0, "video_bitrate": 0, "internet_flow": 0,
         "if_bitrate": 0, "backlog_flow": 1, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 0},
    ],
    "ues_per_slice": [2, 1, 3, 1, 2],
}

QMIX_CFG = {
    # queue-aware enterprise schedulers: backlogged PF | internet flows, alpha 1 | H.264 video, alpha 1 and
    # beta 1 (head-of-line delay in the metric) | backlogged MT; one bearer per UE
    "slices": [
        {"n_slices": 1, "weight": 0.25, "video_app": 0, "video_bitrate": [], "internet_flow": 0, "if_bitrate": [],
         "backlog_flow": 1, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 1},
        {"n_slices": 2, "weight": 0.2, "video_app": 0, "video_bitrate": [], "internet_flow": 1, "if_bitrate": [12],
         "backlog_flow": 0, "algo_alpha": 1, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 1},
        {"n_slices": 2, "weight": 0.1, "video_app": 1, "video_bitrate": [1280], "internet_flow": 0, "if_bitrate": [],
         "backlog_flow": 0, "algo_alpha": 1, "algo_beta": 1, "algo_epsilon": 1, "algo_psi": 1},
        {"n_slices": 1, "weight": 0.15, "video_app": 0, "video_bitrate": [], "internet_flow": 0, "if_bitrate": [],
         "backlog_flow": 1, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 0},
    ],
    "ues_per_slice": [3, 4, 2, 5, 3, 2],
}

QIF_CFG = {
    # like QMIX_CFG without the video application (its trace files do not travel to the GPU box): internet flows under
    # alpha 1 and under alpha 1 + beta 1, for the in-simulator drop-in test
    "slices": [
        {"n_slices": 1, "weight": 0.25, "video_app": 0, "video_bitrate": [], "internet_flow": 0, "if_bitrate": [],
         "backlog_flow": 1, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 1},
        {"n_slices": 2, "weight": 0.2, "video_app": 0, "video_bitrate": [], "internet_flow": 1, "if_bitrate": [12],
         "backlog_flow": 0, "algo_alpha": 1, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 1},
        {"n_slices": 2, "weight": 0.1, "video_app": 0, "video_bitrate": [], "internet_flow": 1, "if_bitrate": [9],
         "backlog_flow": 0, "algo_alpha": 1, "algo_beta": 1, "algo_epsilon": 1, "algo_psi": 1},
        {"n_slices": 1, "weight": 0.15, "video_app": 0, "video_bitrate": [], "internet_flow": 0, "if_bitrate": [],
         "backlog_flow": 1, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1, "algo_psi": 0},
    ],
    "ues_per_slice": [3, 4, 2, 5, 3, 2],
}

FIX20X5 = f"{REF_EXP}/exp-fix20slices/5ues/config-pf.json"
DIFFW = f"{REF_EXP}/exp-customization/exp-backlogged-20slicesdiffw/config.json"
MIX20 = f"{REF_EXP}/exp-customization/exp-backlogged-20slices/config.json"

# name, algo, config, n_ttis, cqi source ("trace" = the simulator's own cqi-traces-noise0 ingest,
# "synth" = injected per-TTI synthetic CQI), seed
CASES = [
    ("a9_fix20x5_trace", 9, FIX20X5, 100, "trace", 1),
    ("a8_fix20x5_trace", 8, FIX20X5, 60, "trace", 1),
    ("a7_fix20x5_trace", 7, FIX20X5, 60, "trace", 1),
    ("a1_fix20x5_trace", 1, FIX20X5, 60, "trace", 1),
    ("a9_fix20x5_synth", 9, FIX20X5, 80, "synth", 2),
    ("a8_fix20x5_synth", 8, FIX20X5, 40, "synth", 3),
    ("a7_fix20x5_synth", 7, FIX20X5, 40, "synth", 4),
    ("a1_fix20x5_synth", 1, FIX20X5, 40, "synth", 5),
    ("a9_diffw_synth", 9, DIFFW, 40, "synth", 6),
    # BASELINE.json configs[0]: SingleCellWithI, scheduler 9, exp-backlogged-20slicesdiffw, real CQI traces
    # through mapping1.config (204 UEs)
    ("a9_diffw_trace", 9, DIFFW, 45, "trace", 1),
    ("a8_diffw_synth", 8, DIFFW, 24, "synth", 7),
    ("a7_mix20_synth", 7, MIX20, 40, "synth", 8),
    ("a9_small_synth", 9, "SMALL", 120, "synth", 9),
    ("a8_small_synth", 8, "SMALL", 120, "synth", 10),
    ("a7_small_synth", 7, "SMALL", 120, "synth", 11),
    ("a1_small_synth", 1, "SMALL", 120, "synth", 12),
    # id 11 (NVS non-greedy, downlink-nvs-scheduler.cpp:405-528): the records also hold every rand() value the
    # 300-sample search drew (rand_ng), read back through ref_harness --rand-log
    ("a11_fix20x5_synth", 11, FIX20X5, 24, "synth", 13),
    ("a11_small_synth", 11, "SMALL", 60, "synth", 14),
    ("a11_fix20x5_trace", 11, FIX20X5, 12, "trace", 1),
    # id 10 (UpperBound, downlink-transport-scheduler.cpp:223-246): an RBG may go to several slices, so the
    # records also hold every (user, RBG) grant in the order of the users' RB lists (ref_harness --alloc-log)
    ("a10_fix20x5_synth", 10, FIX20X5, 40, "synth", 15),
    ("a10_diffw_synth", 10, DIFFW, 24, "synth", 16),
    ("a10_small_synth", 10, "SMALL", 80, "synth", 17),
    ("a10_fix20x5_trace", 10, FIX20X5, 30, "trace", 1),
    # finite queues and head-of-line delays: the records also hold every bearer's dataToTransmit and HoL delay per TTI
    # (ref_harness --queue-log)
    ("a9_qmix_synth", 9, "QMIX", 160, "synth", 21),
    ("a8_qmix_synth", 8, "QMIX", 160, "synth", 22),
    ("a7_qmix_synth", 7, "QMIX", 240, "synth", 23),
    ("a1_qmix_synth", 1, "QMIX", 160, "synth", 24),
    ("a10_qmix_synth", 10, "QMIX", 120, "synth", 25),
    ("a11_qmix_synth", 11, "QMIX", 60, "synth", 26),
    # SubOpt (101) and VogelApproximate (103): DownlinkTransportScheduler(config, 1 / 3), no scenario id
    ("a101_fix20x5_synth", 101, FIX20X5, 50, "synth", 41),
    ("a101_diffw_synth", 101, DIFFW, 24, "synth", 42),
    ("a101_small_synth", 101, "SMALL", 100, "synth", 43),
    ("a103_fix20x5_synth", 103, FIX20X5, 50, "synth", 44),
    ("a103_diffw_synth", 103, DIFFW, 24, "synth", 45),
    ("a103_small_synth", 103, "SMALL", 100, "synth", 46),
    ("a9_qif_synth", 9, "QIF", 160, "synth", 31),
    ("a8_qif_synth", 8, "QIF", 120, "synth", 32),
    ("a7_qif_synth", 7, "QIF", 200, "synth", 33),
    ("a1_qif_synth", 1, "QIF", 120, "synth", 34),
    ("a10_qif_synth", 10, "QIF", 120, "synth", 35),
    ("a101_qif_synth", 101, "QIF", 120, "synth", 36),
    ("a103_qif_synth", 103, "QIF", 120, "synth", 37),
    ("a11_qif_synth", 11, "QIF", 60, "synth", 38),
]


def run_case(name, algo, config, n_ttis, source, seed, tmp):
    queue_aware = config in ("QMIX", "QIF")
    if config == "QIF":
        config = os.path.join(tmp, "qif.json")
        json.dump(QIF_CFG, open(config, "w"))
    if config == "SMALL":
        config = os.path.join(tmp, "small.json")
        json.dump(SMALL_CFG, open(config, "w"))
    if config == "QMIX":
        config = os.path.join(tmp, "qmix.json")
        json.dump(QMIX_CFG, open(config, "w"))
    cfg = json.load(open(config))
    n_ues = int(sum(cfg["ues_per_slice"]))
    n_slices = len(cfg["ues_per_slice"])
    rec_path = os.path.join(tmp, name + ".bin")
    cmd = [HARNESS, "--algo", str(algo), "--config", config, "--ttis", str(n_ttis), "--out", rec_path,
           "--seed", str(seed)]
    rand2 = workload.synth_rand2(seed, 0, 1, 0, n_ttis, n_slices)[:, 0, :]
    rand_path = os.path.join(tmp, name + ".rand")
    rand2.astype("<i4").tofile(rand_path)
    cmd += ["--rand", rand_path]
    if source == "synth":
        cqi = workload.synth_cqi(seed, 0, 1, 0, n_ttis, n_ues, 64)[:, 0]
        cqi_path = os.path.join(tmp, name + ".cqi")
        cqi.tofile(cqi_path)
        cmd += ["--cqi", cqi_path]
    rlog_path = os.path.join(tmp, name + ".rlog")
    if algo == 11:
        cmd += ["--rand-log", rlog_path]
    qlog_path = os.path.join(tmp, name + ".qlog")
    if queue_aware:
        cmd += ["--queue-log", qlog_path]
    alog_path = os.path.join(tmp, name + ".alog")
    if algo == 10:
        cmd += ["--alloc-log", alog_path]
    out = subprocess.run(cmd, check=True, capture_output=True, text=True).stdout.strip().splitlines()[-1]
    rec = golden_io.compact(golden_io.parse_record_stream(rec_path))
    assert rec["T"] == n_ttis, (name, rec["T"])
    if algo == 11:
        raw = np.fromfile(rlog_path, dtype="<i4")
        draws, pos = [], 0
        for _ in range(n_ttis):
            k = int(raw[pos])
            draws.append(raw[pos + 1:pos + 1 + k])
            pos += 1 + k
        assert pos == len(raw)
        width = max(len(d) for d in draws)
        rec["rand_ng_n"] = np.array([len(d) for d in draws], dtype=np.int32)
        rec["rand_ng"] = np.stack([np.pad(d, (0, width - len(d))) for d in draws]).astype(np.int32)
    if queue_aware:
        raw = np.fromfile(qlog_path, dtype=np.uint8).reshape(n_ttis, n_ues * 12)
        rec["queue"] = raw[:, :n_ues * 4].copy().view("<i4").reshape(n_ttis, n_ues)
        rec["hol"] = raw[:, n_ues * 4:].copy().view("<f8").reshape(n_ttis, n_ues)
    if algo == 10:
        raw = np.fromfile(alog_path, dtype="<i2")
        G = int(rec["G"])
        ue = np.full((n_ttis, 2 * G), -1, dtype=np.int16)
        rb = np.full((n_ttis, 2 * G), -1, dtype=np.int16)
        cnt = np.zeros(n_ttis, dtype=np.int32)
        pos = 0
        for t in range(n_ttis):
            k = int(raw[pos]) | (int(raw[pos + 1]) << 16)
            pos += 2
            pairs = raw[pos:pos + 2 * k].reshape(k, 2)
            pos += 2 * k
            cnt[t] = k
            ue[t, :k], rb[t, :k] = pairs[:, 0], pairs[:, 1]
        assert pos == len(raw) and cnt.max() <= 2 * G
        rec["alloc_n"], rec["alloc_ue"], rec["alloc_rbg"] = cnt, ue, rb
    if algo in (8, 9, 10, 101, 103):
        assert (rec["rand2"] == rand2).all(), "scripted rand() values were not the ones consumed"
    rec["config_json"] = json.dumps(cfg)
    rec["source"] = source
    rec["seed"] = seed
    golden_io.save_npz(os.path.join(GOLDEN, name + ".npz"), rec)
    print(f"{name}: {out} -> {os.path.getsize(os.path.join(GOLDEN, name + '.npz')) / 1024:.0f} KiB", flush=True)


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    if not os.path.exists(HARNESS):
        print("oracle/_ref/ref_harness missing: run `make -C oracle ref` first", file=sys.stderr)
        return 1
    os.makedirs(GOLDEN, exist_ok=True)
    with tempfile.TemporaryDirectory() as tmp:
        for case in CASES:
            if args.only and case[0] != args.only:
                continue
            run_case(*case, tmp)
    return 0


if __name__ == "__main__":
    sys.exit(main())
