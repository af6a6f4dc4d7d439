#!/bin/bash
# round 2, GPU session 8: VogelApproximate rewrite (tests + speed), e2e chunk size of the headline leg
O=gpurun_out/s8
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_log_writer.py tests/test_dropin_gpu.py tests/test_two_bearers_gpu.py tests/test_rs_batch_gpu.py -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --ids 103,101,10,11 2>>$O/sweep.err > $O/sweep_ids.jsonl
timeout 900 python tools/fuzz_parity.py --seconds 150 --seed 7 > $O/fuzz.log 2>&1; echo "fuzz rc=$?" >> $O/fuzz.log
for tpl in 8 10 20; do timeout 600 python bench.py --no-cpu-baseline --no-parity-spot --steps 6 --e2e-refresh-ttis-per-launch $tpl --e2e-steps 12 2>>$O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print($tpl, d['value'], d['e2e']['value'], {k:v['value'] for k,v in d['e2e']['variants'].items()})" >> $O/e2e_tpl.txt; done
tail -3 $O/pytest.log; tail -3 $O/fuzz.log; python - <<'PY'
import json
for l in open('gpurun_out/s8/sweep_ids.jsonl'):
    d=json.loads(l); print(d['label'], round(d['cell_ttis_per_s']/1e6,3))
PY
cat $O/e2e_tpl.txt
