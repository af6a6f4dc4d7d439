#!/bin/bash
# round 2, GPU session 9: Vogel v3 (tests, fuzz soak, speed), TTIs per launch
O=gpurun_out/s9
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_log_writer.py tests/test_dropin_gpu.py tests/test_host_api_gpu.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --ids 103,101 2>>$O/sweep.err > $O/sweep_ids.jsonl
timeout 900 python tools/fuzz_parity.py --seconds 200 --seed 7 > $O/fuzz.log 2>&1; echo "fuzz rc=$?" >> $O/fuzz.log
for cfg in "48 16" "96 24" "96 32" "96 48" "128 64"; do set -- $cfg
  timeout 300 python bench.py --kernel-only --steps 8 --warmup 3 --no-parity-spot --ttis-per-step $1 --ttis-per-launch $2 2>>$O/bench_ko.err | sed "s/^/tpl$2 /" >> $O/bench_tpl.jsonl
done
timeout 300 python bench.py --kernel-only --steps 8 --warmup 3 --no-parity-spot --cells 4736 2>>$O/bench_ko.err | sed "s/^/cells4736 /" >> $O/bench_tpl.jsonl
timeout 300 python bench.py --kernel-only --steps 4 --warmup 3 --no-parity-spot --cells 32768 2>>$O/bench_ko.err | sed "s/^/cells32768 /" >> $O/bench_tpl.jsonl
tail -3 $O/pytest.log; tail -3 $O/fuzz.log; python - <<'PY'
import json
for l in open('gpurun_out/s9/sweep_ids.jsonl'):
    d=json.loads(l); print(d['label'], round(d['cell_ttis_per_s']/1e6,3))
for l in open('gpurun_out/s9/bench_tpl.jsonl'):
    v,j=l.split(' ',1); d=json.loads(j); print(v, round(d['value']/1e6,3))
PY
