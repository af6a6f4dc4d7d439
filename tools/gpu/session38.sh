#!/bin/bash
# round 2, closing session: full suite, smoke, bench (both arms), launch list and full capture of the final build
O=gpurun_out/s38
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2>$O/bench.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference.json 2>$O/bench_reference.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/launches_raw.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-spot > $O/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_tti_kernel -s 40 -c 1 -o $O/r02_full_final python bench.py --kernel-only --steps 3 --warmup 3 --no-parity-spot > $O/ncu_full.log 2>&1
timeout 900 python tools/sweep_bench.py --only ids 2>>$O/sweep.err > $O/sweep_ids.jsonl
timeout 900 python tools/sweep_bench.py --only ids --layout 2 2>>$O/sweep.err > $O/sweep_ids_packed.jsonl
timeout 1500 python tools/sweep_bench.py --only sweep 2>>$O/sweep.err > $O/sweep_grid.jsonl
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s38/bench.json').readline())
print('bench', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), {k: round(v['value']/1e6,2) for k,v in d['e2e']['variants'].items()}, 'parity', d['parity_spot']['mismatches'], 'frac', round(d['roofline']['frac'],4))
print(open('gpurun_out/s38/bench_reference.json').readline()[:200])
PY
