#!/bin/bash
# round 2, GPU session 12: number of part-batches (1..4), ncu full capture of a part-batch launch
O=gpurun_out/s12
mkdir -p $O
timeout 600 python -m pytest tests/test_host_api_gpu.py tests/test_gpu_parity.py -m gpu -q -k "half_batches or full_size or async or run_host or generators" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for P in 1 2 3 4; do for rep in 1 2; do
  RS_PARTS=$P timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --no-parity-spot 2>>$O/bench_ko.err | sed "s/^/parts$P /" >> $O/bench_parts.jsonl
done; done
for P in 2 4; do RS_PARTS=$P timeout 600 python bench.py --no-cpu-baseline --no-parity-spot --steps 8 2>>$O/bench.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('parts$P', d['value'], d['e2e']['value'], {k:v['value'] for k,v in d['e2e']['variants'].items()})" >> $O/e2e_parts.txt; done
RS_PARTS=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_tti_kernel -s 40 -c 1 -o $O/r02_full_half python bench.py --kernel-only --steps 3 --warmup 3 --no-parity-spot > $O/ncu_full.log 2>&1
tail -3 $O/pytest.log; python - <<'PY'
import json
for l in open('gpurun_out/s12/bench_parts.jsonl'):
    v,j=l.split(' ',1); d=json.loads(j); print(v, round(d['value']/1e6,3))
PY
cat $O/e2e_parts.txt
