#!/bin/bash
# round 2, GPU session 2: full GPU suite (12-s record, plug-in on rs_step_cell), properly timed bench, drop-in latency,
# reference baselines, ncu captures
O=gpurun_out/s2
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for v in product nobulk; do
  lib=$PWD/build/librs_$v.so; [ $v = product ] && lib=$PWD/radiosaber_b200/librs_sched.so
  for rep in 1 2 3; do
    RS_SCHED_LIB=$lib timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --no-parity-spot 2>>$O/bench_ko.err | sed "s/^/$v /" >> $O/bench_ko.jsonl
  done
done
timeout 900 python bench.py > $O/bench_product.json 2> $O/bench_product.err; echo "bench rc=$?" >> $O/bench_product.err
timeout 600 python tools/dropin_latency.py > $O/dropin_latency.jsonl 2> $O/dropin_latency.err
for opt in O0 O2; do for reg in all alloc; do
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 --ref-opt $opt --ref-region $reg >> $O/reference_arm.jsonl 2>> $O/reference_arm.err
done; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_raw.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $O/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rs_tti_kernel -s 3 -c 1 -o $O/r02_full python bench.py --kernel-only --steps 3 --warmup 3 --no-parity-spot > $O/ncu_full.log 2>&1
tail -3 $O/pytest.log; cut -c1-160 $O/bench_ko.jsonl; cat $O/dropin_latency.jsonl; cut -c1-400 $O/reference_arm.jsonl
