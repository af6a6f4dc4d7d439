#!/bin/bash
# sort with scan units on the wide (512-thread) kernel: big-cell sweep points, units build vs previous build
O=gpurun_out/s21
mkdir -p $O
P="50,40;20,40;40,20;30,30;50,10;30,20"
RS_SCHED_LIB=$PWD/build/librs_units.so timeout 600 python tools/sweep_bench.py --only sweep --points "$P" > $O/units.jsonl 2>$O/units.err
RS_SCHED_LIB=$PWD/build/librs_old.so timeout 600 python tools/sweep_bench.py --only sweep --points "$P" > $O/old.jsonl 2>$O/old.err
python - <<'PY'
import json
for f in ("units","old"):
    for l in open(f"gpurun_out/s21/{f}.jsonl"):
        d=json.loads(l); print(f, d["label"], round(d["cell_ttis_per_s"]/1e6,3), d["smem_bytes_per_cta"])
PY
tail -3 $O/units.err
