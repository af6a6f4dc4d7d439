#!/bin/bash
# 10 cells per SM (48-register cap) against 9 on the packed CQI layout (19.8 KB of shared memory per cell)
O=gpurun_out/s24
mkdir -p $O
for lib in mb10 cur; do
  [ $lib = mb10 ] && export RS_SCHED_LIB=$PWD/build/librs_mb10.so || unset RS_SCHED_LIB
  timeout 600 python tools/sweep_bench.py --only ids --ids 9,8,103 --layout 2 --launches 6 2>>$O/err > $O/$lib.jsonl
  timeout 600 python tools/sweep_bench.py --only ids --ids 9 --layout 0 --launches 6 2>>$O/err >> $O/$lib.jsonl
done
python - <<'PY'
import json
for f in ("mb10","cur"):
    for l in open(f"gpurun_out/s24/{f}.jsonl"):
        d=json.loads(l); print(f, d["label"], "layout", d["cqi_layout"], round(d["cell_ttis_per_s"]/1e6,3), d["smem_bytes_per_cta"])
PY
tail -3 $O/err
