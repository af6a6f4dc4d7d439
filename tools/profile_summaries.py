#!/usr/bin/env python3
"""Turn the ncu artefacts a gpurun call leaves in gpurun_out/ into the summaries kept under profiles/.

  profile_summaries.py launches <launches_raw.csv> <out.csv>   per-kernel share of the launch list
  profile_summaries.py full <report.ncu-rep> <out.csv>         selected raw metrics of one kernel launch
"""
import csv
import re
import subprocess
import sys

KEEP = re.compile(
    r"^(dram__bytes_(read|write)\.sum|gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed|gpu__time_duration\.sum|"
    r"l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum|l1tex__data_pipe_lsu_wavefronts_mem_shared\.sum|"
    r"l1tex__t_sector_hit_rate\.pct|lts__t_sector_hit_rate\.pct|"
    r"launch__(block_size|grid_size|occupancy_limit_(registers|shared_mem|warps|blocks)|registers_per_thread|"
    r"shared_mem_per_block_dynamic|waves_per_multiprocessor|sm_count)|"
    r"sm__inst_executed_pipe_(alu|fp64|lsu|xu|fma)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_tensor_cycles_active\.avg\.pct_of_peak_sustained_active|sm__throughput\.avg\.pct_of_peak_sustained_elapsed|"
    r"sm__warps_active\.avg\.pct_of_peak_sustained_active|sm__inst_executed\.sum|smsp__inst_executed\.sum|"
    r"sm__inst_executed\.avg\.per_cycle_elapsed|smsp__issue_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__warps_eligible\.avg\.per_cycle_active|smsp__warps_active\.avg\.per_cycle_active|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|sm__cycles_elapsed\.max)$")


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mu = hdr.index("Metric Unit")
    agg = {}
    for r in rows[1:]:
        if len(r) != len(hdr):
            continue
        v = float(r[mv].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
        name = re.sub(r"\(.*", "", r[kn]).strip()
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write("# per-kernel aggregate of the ncu launch list (gpu__time_duration.sum, --clock-control none);\n")
        f.write("# per-launch times under the profiler are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write("kernel,launches,total_ms,share,mean_ms\n")
        for name, (n, ms) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"{name},{n},{ms:.3f},{ms / tot:.4f},{ms / n:.4f}\n")


def full(rep, dst):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(dst, "w") as f:
        f.write("metric,unit,value\n")
        for h, u, v in sorted(zip(hdr, units, vals)):
            if KEEP.match(h):
                f.write(f"{h},{u},{v.replace(',', '')}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
