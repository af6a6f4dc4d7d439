#!/bin/bash
# round 2, GPU session 10: two-stream half-batches (tests, A/B), full suite, final bench line
O=gpurun_out/s10
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for v in split nosplit; do
  if [ $v = nosplit ]; then export RS_NO_SPLIT=1; else unset RS_NO_SPLIT; fi
  for rep in 1 2; do timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --no-parity-spot 2>>$O/bench_ko.err | sed "s/^/$v /" >> $O/bench_ko.jsonl; done
  timeout 300 python bench.py --kernel-only --steps 8 --warmup 3 --no-parity-spot --ttis-per-step 96 --ttis-per-launch 16 2>>$O/bench_ko.err | sed "s/^/${v}_96 /" >> $O/bench_ko.jsonl
  timeout 300 python bench.py --kernel-only --steps 8 --warmup 3 --no-parity-spot --ttis-per-step 96 --ttis-per-launch 8 2>>$O/bench_ko.err | sed "s/^/${v}_96_8 /" >> $O/bench_ko.jsonl
done
unset RS_NO_SPLIT
timeout 900 python bench.py > $O/bench_product.json 2> $O/bench_product.err; echo "bench rc=$?" >> $O/bench_product.err
tail -3 $O/pytest.log; python - <<'PY'
import json
for l in open('gpurun_out/s10/bench_ko.jsonl'):
    v,j=l.split(' ',1); d=json.loads(j); print(v, round(d['value']/1e6,3))
d=json.load(open('gpurun_out/s10/bench_product.json'))
print('bench', d['value'], d['parity_spot']['mismatches'], d['e2e']['value'], {k:v['value'] for k,v in d['e2e']['variants'].items()})
PY
