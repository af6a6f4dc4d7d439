#!/bin/bash
# id 11: the 300-sample search on RBG bitmasks: tests and speed
O=gpurun_out/s33
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -x -k "103 or fuzz or fixed or nongreedy_slice" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --ids 103,101 2>>$O/err > $O/ids.jsonl
timeout 600 python tools/sweep_bench.py --only ids --ids 103 --layout 2 2>>$O/err >> $O/ids.jsonl
RS_SCHED_LIB=$PWD/build/librs_old.so timeout 600 python tools/sweep_bench.py --only ids --ids 103 2>>$O/err > $O/ids_old.jsonl
python - <<'PY'
import json
for f in ("ids","ids_old"):
    for l in open(f"gpurun_out/s33/{f}.jsonl"):
        d=json.loads(l)
        if "mix" in d["label"]: continue
        print(f, d["label"], d.get("cqi_layout"), round(d["cell_ttis_per_s"]/1e6,3))
PY
