#!/usr/bin/env python3
"""Per-TTI cost of the in-simulator drop-in at B = 1: the same LTE-Sim scenario once with the
reference's own scheduler (oracle/_ref/ref_harness) and once with RsGpuScheduler installed in the eNB
(oracle/_ref/ref_harness_gpu --gpu: host plug-in -> rs_step_cell -> CUDA), timed inside the harness around
DoSchedule, with DoStopSchedule (RLC, packets, log lines: the same code in both) timed apart.  Needs a GPU for the
second arm.  Prints one JSON line per scheduler id."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from radiosaber_b200 import workload  # noqa: E402

S, U, T = 20, 100, int(os.environ.get("RS_LAT_TTIS", "400"))


def run(binary, algo, tmp, gpu):
    cmd = [binary, "--algo", str(algo), "--config", os.path.join(tmp, "cfg.json"), "--ttis", str(T), "--seed", "1",
           "--rand", os.path.join(tmp, "r.rand"), "--cqi", os.path.join(tmp, "c.cqi"), "--time-every", str(T // 4)]
    if gpu:
        cmd.append("--gpu")
    out = subprocess.run(cmd, capture_output=True, text=True, cwd=tmp)
    marks = [json.loads(l) for l in out.stdout.splitlines() if l.startswith('{"sched_calls"')]
    if len(marks) < 4:
        return None
    # skip the first quarter (context creation, first launches)
    n = marks[-1]["sched_calls"] - marks[0]["sched_calls"]
    tot = marks[-1]["sched_seconds"] - marks[0]["sched_seconds"]
    stop = marks[-1]["stop_seconds"] - marks[0]["stop_seconds"]
    res = {"do_schedule_ms": 1e3 * tot / n, "do_stop_schedule_ms": 1e3 * stop / n, "ewma_select_alloc_ms": 1e3 * (tot - stop) / n}
    if gpu:
        res["rs_step_cell_ms"] = 1e3 * (marks[-1].get("abi_seconds", 0.0) - marks[0].get("abi_seconds", 0.0)) / n
    return res


def main():
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_harness")
    ref_o2 = os.path.join(ROOT, "oracle", "_ref", "O2", "ref_harness")
    gpu = os.path.join(ROOT, "oracle", "_ref", "ref_harness_gpu")
    gpu_o2 = os.path.join(ROOT, "oracle", "_ref", "O2", "ref_harness_gpu")
    with tempfile.TemporaryDirectory() as tmp:
        json.dump({"slices": [{"n_slices": S, "weight": 1.0 / S, "video_app": 0, "video_bitrate": 0, "internet_flow": 0,
                               "if_bitrate": 0, "backlog_flow": 1, "algo_alpha": 0, "algo_beta": 0, "algo_epsilon": 1,
                               "algo_psi": 1}], "ues_per_slice": [U // S] * S}, open(os.path.join(tmp, "cfg.json"), "w"))
        workload.synth_rand2(1, 0, 1, 0, T, S)[:, 0, :].astype("<i4").tofile(os.path.join(tmp, "r.rand"))
        workload.synth_cqi(1, 0, 1, 0, T, U, 64)[:, 0].tofile(os.path.join(tmp, "c.cqi"))
        for algo in (9, 8, 7, 1):
            line = {"scheduler_id": algo, "cell": "20 slices x 5 UEs, 100 MHz, backlogged", "ttis": T,
                    "reference_O0": run(ref, algo, tmp, False),
                    "reference_O2": run(ref_o2, algo, tmp, False) if os.path.exists(ref_o2) else None,
                    "plugin_gpu_hostO0": run(gpu, algo, tmp, True),
                    "plugin_gpu_hostO2": run(gpu_o2, algo, tmp, True) if os.path.exists(gpu_o2) else None}
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
