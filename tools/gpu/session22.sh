#!/bin/bash
# id 103: one walking warp + packed candidates: tests and speed
O=gpurun_out/s22
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_dropin_gpu.py -m gpu -q -x -k "103 or fuzz or sort" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --ids 103,101,10 2>>$O/sweep.err > $O/sweep_ids.jsonl
RS_SCHED_LIB=$PWD/build/librs_old.so timeout 600 python tools/sweep_bench.py --only ids --ids 103 2>>$O/sweep.err > $O/sweep_ids_old.jsonl
python - <<'PY'
import json
for f in ("sweep_ids","sweep_ids_old"):
    for l in open(f"gpurun_out/s22/{f}.jsonl"):
        d=json.loads(l); print(f, d["label"], round(d["cell_ttis_per_s"]/1e6,3))
PY
