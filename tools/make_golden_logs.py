#!/usr/bin/env python3
"""Generate tests/golden/logs/<case>.{stdout,stderr}: the text the UNMODIFIED reference prints for the
first TTIS TTIs of a golden case (same harness, inputs and seeds as tools/make_golden.py, so TTI t of the
text belongs to TTI t of tests/golden/<case>.npz).  stdout = the allocation dump of RBsAllocation
(downlink-transport-scheduler.cpp:523-527, 631-649; downlink-nvs-scheduler.cpp:314-332), stderr = the
all_bytes line (:374) and the cumu_bytes / cumu_rbs lines the paper's plotters parse (:192-199,
NSDI23-radiosaber-experiments/exp-customization/plot_throughput.py:35-47).

Usage: python tools/make_golden_logs.py   (needs oracle/_ref/ref_harness, i.e. /root/reference mounted)
"""
from __future__ import annotations

import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosaber_b200 import workload  # noqa: E402
from tools import make_golden as mg  # noqa: E402

TTIS = 6
CASES = ["a9_fix20x5_synth", "a8_fix20x5_synth", "a7_fix20x5_synth", "a1_fix20x5_synth", "a9_small_synth",
         "a9_fix20x5_trace", "a9_qmix_synth", "a7_qmix_synth", "a1_qmix_synth",
         "a10_fix20x5_synth", "a10_qmix_synth", "a11_fix20x5_synth", "a101_fix20x5_synth", "a103_fix20x5_synth"]
TTIS_OF = {"a9_qmix_synth": 40, "a7_qmix_synth": 60, "a1_qmix_synth": 40, "a10_qmix_synth": 40}   # long enough for the finite flows to show up
OUT = os.path.join(ROOT, "tests", "golden", "logs")


def main():
    os.makedirs(OUT, exist_ok=True)
    table = {c[0]: c for c in mg.CASES}
    with tempfile.TemporaryDirectory() as tmp:
        for name in CASES:
            _, algo, config, _, source, seed = table[name]
            if config == "SMALL":
                config = os.path.join(tmp, "small.json")
                json.dump(mg.SMALL_CFG, open(config, "w"))
            if config == "QMIX":
                config = os.path.join(tmp, "qmix.json")
                json.dump(mg.QMIX_CFG, open(config, "w"))
            ttis = TTIS_OF.get(name, TTIS)
            cfg = json.load(open(config))
            n_ues, n_slices = int(sum(cfg["ues_per_slice"])), len(cfg["ues_per_slice"])
            cmd = [mg.HARNESS, "--algo", str(algo), "--config", config, "--ttis", str(ttis), "--seed", str(seed),
                   "--log-out", os.path.join(OUT, name)]
            rand_path = os.path.join(tmp, name + ".rand")
            workload.synth_rand2(seed, 0, 1, 0, ttis, n_slices)[:, 0, :].astype("<i4").tofile(rand_path)
            cmd += ["--rand", rand_path]
            if source == "synth":
                cqi_path = os.path.join(tmp, name + ".cqi")
                workload.synth_cqi(seed, 0, 1, 0, ttis, n_ues, 64)[:, 0].tofile(cqi_path)
                cmd += ["--cqi", cqi_path]
            subprocess.run(cmd, check=True, capture_output=True, text=True)
            for ext in ("stdout", "stderr"):
                p = os.path.join(OUT, f"{name}.{ext}")
                print(f"{p}: {os.path.getsize(p)} bytes")


if __name__ == "__main__":
    main()
