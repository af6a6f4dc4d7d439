#!/bin/bash
# round 2, GPU session 4: full suite after the two-bearer / FixedShape / Dims changes, kernel A/B on one box, phase split of big cells
O=gpurun_out/s4
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for v in product nofixed s2 nobulk nobulk_nofixed; do
  lib=$PWD/radiosaber_b200/librs_sched.so; case $v in nobulk*) lib=$PWD/build/librs_nobulk.so;; s2) lib=$PWD/build/librs_s2.so;; esac
  nf=""; case $v in *nofixed) nf=1;; esac
  for rep in 1 2; do
    if [ -n "$nf" ]; then export RS_NO_FIXED_SHAPE=1; else unset RS_NO_FIXED_SHAPE; fi
    RS_SCHED_LIB=$lib timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --no-parity-spot 2>>$O/bench_ko.err | sed "s/^/$v /" >> $O/bench_ko.jsonl
  done
done
unset RS_NO_FIXED_SHAPE
timeout 600 python tools/dropin_latency.py > $O/dropin_latency.jsonl 2> $O/dropin_latency.err
RS_SCHED_LIB=$PWD/build/librs_phase.so timeout 600 python tools/sweep_bench.py --only sweep --points "20,5;20,40;50,40;50,10" --launches 2 > $O/phase_sweep.txt 2>&1
timeout 900 python bench.py > $O/bench_product.json 2> $O/bench_product.err; echo "bench rc=$?" >> $O/bench_product.err
tail -3 $O/pytest.log; cut -c1-150 $O/bench_ko.jsonl; cat $O/dropin_latency.jsonl; grep "cta " $O/phase_sweep.txt | head -40
