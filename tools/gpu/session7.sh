#!/bin/bash
# round 2, GPU session 7: full suite, full bench line, configs[2]/[3] sweeps, staging-rule A/B, ncu captures of the final kernel
O=gpurun_out/s7
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py > $O/bench_product.json 2> $O/bench_product.err; echo "bench rc=$?" >> $O/bench_product.err
timeout 1500 python tools/sweep_bench.py --out $O/configs2_3_sweep.jsonl > /dev/null 2>>$O/sweep.err
PTS="15,15;20,10;30,10;15,10;10,15;30,5;40,5"
timeout 600 python tools/sweep_bench.py --only sweep --points "$PTS" 2>>$O/sweep.err | sed "s/^/fit8 /" >> $O/sweep_stagefit_ab.jsonl
RS_STAGE_MIN_FIT=6 timeout 600 python tools/sweep_bench.py --only sweep --points "$PTS" 2>>$O/sweep.err | sed "s/^/fit6 /" >> $O/sweep_stagefit_ab.jsonl
RS_STAGE_MIN_FIT=4 timeout 600 python tools/sweep_bench.py --only sweep --points "$PTS" 2>>$O/sweep.err | sed "s/^/fit4 /" >> $O/sweep_stagefit_ab.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_raw.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-spot > $O/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_tti_kernel -s 3 -c 1 -o $O/r02_full_fixed python bench.py --kernel-only --steps 3 --warmup 3 --no-parity-spot > $O/ncu_full.log 2>&1
tail -3 $O/pytest.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/s7/bench_product.json'))
print('bench', d['value'], d['parity_spot']['mismatches'], d['e2e']['value'], {k:v['value'] for k,v in d['e2e']['variants'].items()})
for l in open('gpurun_out/s7/configs2_3_sweep.jsonl'):
    d=json.loads(l); print(d['label'], round(d['cell_ttis_per_s']/1e6,3), d['smem_bytes_per_cta'])
for l in open('gpurun_out/s7/sweep_stagefit_ab.jsonl'):
    v,j=l.split(' ',1); d=json.loads(j); print(v, d['label'], round(d['cell_ttis_per_s']/1e6,3), d['smem_bytes_per_cta'])
PY
