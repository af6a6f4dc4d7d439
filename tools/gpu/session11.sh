#!/bin/bash
# round 2, GPU session 11: full suite, final bench line, configs[2]/[3] sweeps and the ncu launch list of the final build
O=gpurun_out/s11
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python bench.py > $O/bench_product.json 2> $O/bench_product.err; echo "bench rc=$?" >> $O/bench_product.err
timeout 1500 python tools/sweep_bench.py --out $O/configs2_3_sweep.jsonl > /dev/null 2>>$O/sweep.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_raw.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-spot > $O/ncu_launch_bench.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>> $O/bench_reference.err
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" >> $O/smoke.log
tail -3 $O/pytest.log; tail -2 $O/smoke.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/s11/bench_product.json'))
print('bench', d['value'], d['ms_per_step'], d['parity_spot']['mismatches'], d['e2e']['value'], {k:v['value'] for k,v in d['e2e']['variants'].items()}, d['gpu_launches'], d['roofline']['frac'])
for l in open('gpurun_out/s11/configs2_3_sweep.jsonl'):
    d=json.loads(l); print(d['label'], round(d['cell_ttis_per_s']/1e6,3), d['smem_bytes_per_cta'])
PY
