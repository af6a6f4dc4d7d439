#!/usr/bin/env python3
"""Top source lines of an `ncu --import-source on` report by executed warp instructions.
usage: ncu_src_lines.py <report.ncu-rep> [top_n] [divide_by]   (dev tool; `ncu --page source --print-source cuda,sass`)"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40; div = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
agg, hdr = {}, None
for r in csv.reader(txt.splitlines()):
    if "Instructions Executed" in r:
        hdr = r; ci = r.index("Instructions Executed"); cs = r.index("# Samples"); continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        try: n = int(r[ci])
        except ValueError: continue
        a = agg.setdefault(int(r[0]), [0, 0, r[1].strip()[:110]])
        a[0] += n; a[1] += int(r[cs] or 0)
tot = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
print("total warp instructions", tot, "samples", ts)
for ln, (n, s, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"  {100*n/tot:5.1f}% {n/div:9.0f}  samples {100*s/max(ts,1):5.1f}%  L{ln}: {t}")
