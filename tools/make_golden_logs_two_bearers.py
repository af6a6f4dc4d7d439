#!/usr/bin/env python3
"""Generate tests/golden/logs_two_bearers/<id>.{stdout,stderr}: what the UNMODIFIED reference prints for a cell whose
internet-flow slices carry two bearers per UE (tests/data/cfg_two_bearers.json; `internet_flow: 2` as in
NSDI23-radiosaber-experiments/exp-customization/exp-customize-20slices/config.json), with the seeded CQI and rand()
inputs tests/test_dropin_gpu.py feeds the plug-in.  The record format of ref_harness is per UE, so for these runs the
reference's own log text is the golden.  Needs oracle/_ref/ref_harness (i.e. /root/reference mounted)."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosaber_b200 import workload  # noqa: E402

CFG = os.path.join(ROOT, "tests", "data", "cfg_two_bearers.json")
OUT = os.path.join(ROOT, "tests", "golden", "logs_two_bearers")
HARNESS = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
TTIS, SEED = 120, 41
IDS = (9, 8, 7, 10, 101, 103)


def n_bearers(cfg):
    per_slice = [g["video_app"] + g["internet_flow"] + g["backlog_flow"] for g in cfg["slices"] for _ in range(g["n_slices"])]
    return sum(n * b for n, b in zip(cfg["ues_per_slice"], per_slice))


def write_inputs(tmp, cfg):
    U, S = sum(cfg["ues_per_slice"]), len(cfg["ues_per_slice"])
    cqi, rnd = os.path.join(tmp, "cqi.bin"), os.path.join(tmp, "rand.bin")
    workload.synth_cqi(SEED, 0, 1, 0, TTIS, U, 64)[:, 0].tofile(cqi)
    workload.synth_rand2(SEED, 0, 1, 0, TTIS, S)[:, 0, :].astype("<i4").tofile(rnd)
    return cqi, rnd


def command(harness, algo, cqi, rnd, prefix, cfg):
    return [harness, "--algo", str(algo), "--config", CFG, "--ttis", str(TTIS), "--seed", str(SEED), "--cqi", cqi,
            "--rand", rnd, "--bearers", str(n_bearers(cfg)), "--log-out", prefix]


def main():
    os.makedirs(OUT, exist_ok=True)
    cfg = json.load(open(CFG))
    with tempfile.TemporaryDirectory() as tmp:
        cqi, rnd = write_inputs(tmp, cfg)
        for algo in IDS:
            subprocess.run(command(HARNESS, algo, cqi, rnd, os.path.join(OUT, f"a{algo}"), cfg), check=True, capture_output=True)
            err = open(os.path.join(OUT, f"a{algo}.stderr")).read()
            both = 0   # TTIs in which some UE had both of its bearers served
            seen = {}
            for l in err.splitlines():
                f = l.split()
                if len(f) > 10 and f[1] == "app:":
                    k = (f[0], f[10])
                    both += k in seen
                    seen[k] = 1
            print(f"id {algo}: {os.path.getsize(os.path.join(OUT, f'a{algo}.stdout'))} + {len(err)} bytes, "
                  f"{both} (TTI, user) pairs with two bearers served")


if __name__ == "__main__":
    main()
