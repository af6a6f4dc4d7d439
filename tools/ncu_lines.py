#!/usr/bin/env python3
"""Aggregate an ncu source-page CSV (SASS rows) per CUDA source line.

usage: ncu_lines.py <report.ncu-rep> <kernel-substring> [top_n]
Joins `ncu --page source --csv` (per-SASS-instruction counters) with the line table of the cubin
inside radiosaber_b200/librs_sched.so (nvdisasm -g) by instruction order, then prints executed
warp instructions, stall samples and shared-memory wavefronts per source line.  Dev tool.
"""
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_table(kernel_sub):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "radiosaber_b200", "librs_sched.so")],
                              cwd=tmp, stdout=subprocess.DEVNULL)
        cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    lines, cur, on = [], None, False
    for l in dis.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
        if m:
            on = kernel_sub in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = int(m.group(2))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    return lines


def main():
    rep, ksub = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 50
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
    hdr = rows[hdr_i]
    ci, cs, cw = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
    sass = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    lt = line_table(ksub)
    if len(lt) != len(sass):
        print(f"warning: {len(sass)} SASS rows in report vs {len(lt)} in cubin (rebuild mismatch?)", file=sys.stderr)
    src = open(os.path.join(ROOT, "radiosaber_b200", "csrc", "rs_device.cuh")).read().splitlines()
    agg = {}
    tot = tots = 0
    for k, r in enumerate(sass):
        ln = lt[k] if k < len(lt) else None
        n, s, w = int(r[ci]), int(r[cs]), int(r[cw] or 0)
        a = agg.setdefault(ln, [0, 0, 0])
        a[0] += n; a[1] += s; a[2] += w
        tot += n; tots += s
    print(f"total warp instructions {tot}, samples {tots}")
    for ln, (n, s, w) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = src[ln - 1].strip()[:100] if ln and ln <= len(src) else "?"
        print(f"{100 * n / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% stall  smem_wf {w:>11}  L{ln}: {text}")


if __name__ == "__main__":
    main()
