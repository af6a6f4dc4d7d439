#!/bin/bash
# full GPU suite on the current build + ids sweep in both CQI layouts
O=gpurun_out/s25
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --layout 2 2>>$O/err > $O/ids_layout2.jsonl
timeout 600 python tools/sweep_bench.py --only ids 2>>$O/err > $O/ids_layout0.jsonl
python - <<'PY'
import json
for f in ("ids_layout2","ids_layout0"):
    for l in open(f"gpurun_out/s25/{f}.jsonl"):
        d=json.loads(l)
        if "mix" in d["label"]: continue
        print(f, d["label"], round(d["cell_ttis_per_s"]/1e6,3), d["smem_bytes_per_cta"])
PY
