#!/usr/bin/env python3
"""tests/golden/two_bearers/a<id>.{stdout,stderr}.gz: the text the UNMODIFIED reference prints during the very runs
tools/make_golden_two_bearers.py records per bearer (same config, seed, CQI and rand() inputs, TTI count), so that the
log writer can be checked in batch mode -- per-bearer queues, delays and application ids in, the reference's text out
(tests/test_log_writer.py).  Needs oracle/_ref/ref_harness (i.e. /root/reference mounted)."""
import gzip
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosaber_b200 import workload  # noqa: E402
from tools.make_golden_logs_two_bearers import CFG, HARNESS, n_bearers  # noqa: E402
from tools.make_golden_two_bearers import OUT, SEED, TTIS  # noqa: E402

IDS = (9, 8, 7, 10, 101, 103, 11)   # id 1 schedules flows: its two-bearer cell is logged as one user per bearer


def main():
    cfg = json.load(open(CFG))
    U, S = sum(cfg["ues_per_slice"]), len(cfg["ues_per_slice"])
    with tempfile.TemporaryDirectory() as tmp:
        cqi, rnd = os.path.join(tmp, "cqi.bin"), os.path.join(tmp, "rand.bin")
        workload.synth_cqi(SEED, 0, 1, 0, TTIS, U, 64)[:, 0].tofile(cqi)
        workload.synth_rand2(SEED, 0, 1, 0, TTIS, S)[:, 0, :].astype("<i4").tofile(rnd)
        for algo in IDS:
            prefix = os.path.join(tmp, f"a{algo}")
            subprocess.run([HARNESS, "--algo", str(algo), "--config", CFG, "--ttis", str(TTIS), "--seed", str(SEED), "--cqi", cqi,
                            "--rand", rnd, "--bearers", str(n_bearers(cfg)), "--log-out", prefix], check=True, capture_output=True)
            for ext in ("stdout", "stderr"):
                text = open(f"{prefix}.{ext}", "rb").read()
                with open(os.path.join(OUT, f"a{algo}.{ext}.gz"), "wb") as f:
                    f.write(gzip.compress(text, 9, mtime=0))
                print(f"id {algo} {ext}: {len(text)} bytes")


if __name__ == "__main__":
    main()
