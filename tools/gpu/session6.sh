#!/bin/bash
# round 2, GPU session 6: id 10 warp sorts (tests), f4 ids + phase split, headline table vs direct, big cells staged vs not
O=gpurun_out/s6
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_log_writer.py tests/test_dropin_gpu.py tests/test_two_bearers_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 900 python tools/sweep_bench.py --only ids --ids 10,101,103,11,9,8,7,1 2>>$O/sweep.err > $O/sweep_ids.jsonl
RS_SCHED_LIB=$PWD/build/librs_phase.so timeout 600 python tools/sweep_bench.py --only ids --ids 10,101,103,11 --launches 1 > $O/phase_ids.txt 2>&1
for v in nofixed_table nofixed_direct; do
  export RS_NO_FIXED_SHAPE=1; unset RS_FORCE_DIRECT; [ $v = nofixed_direct ] && export RS_FORCE_DIRECT=1
  for rep in 1 2; do timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --no-parity-spot 2>>$O/bench_ko.err | sed "s/^/$v /" >> $O/bench_ko.jsonl; done
done
unset RS_NO_FIXED_SHAPE RS_FORCE_DIRECT
timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --no-parity-spot 2>>$O/bench_ko.err | sed "s/^/product /" >> $O/bench_ko.jsonl
PTS="20,40;40,20;50,40;50,10;20,20"
RS_NO_STAGE=1 timeout 900 python tools/sweep_bench.py --only sweep --points "$PTS" 2>>$O/sweep.err | sed "s/^/nostage /" >> $O/sweep_stage_ab.jsonl
tail -3 $O/pytest.log; python - <<'PY'
import json
for f in ('gpurun_out/s6/sweep_ids.jsonl',):
    for l in open(f):
        d=json.loads(l); print(d['label'], round(d['cell_ttis_per_s']/1e6,3))
for l in open('gpurun_out/s6/bench_ko.jsonl'):
    v,j=l.split(' ',1); d=json.loads(j); print(v, round(d['value']/1e6,3))
for l in open('gpurun_out/s6/sweep_stage_ab.jsonl'):
    v,j=l.split(' ',1); d=json.loads(j); print(v, d['label'], round(d['cell_ttis_per_s']/1e6,3), d['cells'], d['smem_bytes_per_cta'])
PY
grep "^cta 0 \|label" $O/phase_ids.txt | cut -c1-140 | awk '/label/ || NR%8==1' | head -30
