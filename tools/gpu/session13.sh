#!/bin/bash
# round 2, GPU session 13: scan-free Vogel walk (tests, fuzz, speed); sweep of ids with multi-launch calls
O=gpurun_out/s13
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_log_writer.py tests/test_dropin_gpu.py tests/test_two_bearers_gpu.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/fuzz_parity.py --seconds 100 --seed 11 > $O/fuzz.log 2>&1; echo "fuzz rc=$?" >> $O/fuzz.log
timeout 900 python tools/sweep_bench.py --only ids 2>>$O/sweep.err > $O/sweep_ids.jsonl
tail -3 $O/pytest.log; tail -2 $O/fuzz.log; python - <<'PY'
import json
for l in open('gpurun_out/s13/sweep_ids.jsonl'):
    d=json.loads(l); print(d['label'], round(d['cell_ttis_per_s']/1e6,3))
PY
