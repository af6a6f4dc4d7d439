#!/bin/bash
# round 2, GPU session 3: two bearers on the device, FixedShape kernel A/B, bulk vs cp.async on big cells, drop-in latency
O=gpurun_out/s3
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for v in product nofixed nobulk nobulk_nofixed; do
  lib=$PWD/radiosaber_b200/librs_sched.so; case $v in nobulk*) lib=$PWD/build/librs_nobulk.so;; esac
  nf=""; case $v in *nofixed) nf=1;; esac
  for rep in 1 2; do
    RS_NO_FIXED_SHAPE=$nf RS_SCHED_LIB=$lib timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 --no-parity-spot 2>>$O/bench_ko.err | sed "s/^/$v /" >> $O/bench_ko.jsonl
  done
done
sed -i 's/RS_NO_FIXED_SHAPE=/RS_NO_FIXED_SHAPE=/' /dev/null
for v in product nobulk; do
  lib=$PWD/radiosaber_b200/librs_sched.so; [ $v = nobulk ] && lib=$PWD/build/librs_nobulk.so
  RS_SCHED_LIB=$lib timeout 600 python tools/sweep_bench.py --only sweep --points "5,40;20,20;20,40;50,10;50,40;10,10;50,2" 2>>$O/sweep.err | sed "s/^/$v /" >> $O/sweep_bulk_ab.jsonl
done
timeout 600 python tools/dropin_latency.py > $O/dropin_latency.jsonl 2> $O/dropin_latency.err
timeout 900 python tools/sweep_bench.py --only ids --out $O/sweep_ids.jsonl > /dev/null 2>>$O/sweep.err
tail -3 $O/pytest.log; cut -c1-150 $O/bench_ko.jsonl; cut -c1-260 $O/sweep_bulk_ab.jsonl; cat $O/dropin_latency.jsonl
