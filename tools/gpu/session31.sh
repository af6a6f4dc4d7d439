#!/bin/bash
# full GPU suite on the build with den/cnt aliasing and per-instantiation register caps; ids sweep in both layouts; bench line
O=gpurun_out/s31
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --layout 2 2>>$O/err > $O/ids_layout2.jsonl
timeout 600 python tools/sweep_bench.py --only ids 2>>$O/err > $O/ids_layout0.jsonl
python - <<'PY'
import json
for f in ("ids_layout2","ids_layout0"):
    for l in open(f"gpurun_out/s31/{f}.jsonl"):
        d=json.loads(l)
        if "mix" in d["label"]: continue
        print(f, d["label"], round(d["cell_ttis_per_s"]/1e6,3), d["smem_bytes_per_cta"])
PY
timeout 900 python bench.py > $O/bench.json 2>$O/bench.err; cut -c1-400 $O/bench.json
