#!/bin/bash
# round 2, GPU session 1: GPU test suite on the new host pipeline + bulk staging, A/B of the partition scan variants
mkdir -p gpurun_out/s1
O=gpurun_out/s1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/gpu.txt 2>&1
nproc >> $O/gpu.txt
timeout 1500 python -m pytest tests -m gpu -q -k "not long_records_exist" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
for v in twoptr_bulk twoptr; do
  RS_SCHED_LIB=$PWD/build/librs_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_long_horizon.py -m gpu -q -k "sort or headline or golden or reference_record or fuzz or long or sweep" > $O/pytest_$v.log 2>&1; echo "rc=$?" >> $O/pytest_$v.log
done
for v in product nobulk twoptr twoptr_bulk; do
  lib=$PWD/build/librs_$v.so; [ $v = product ] && lib=$PWD/radiosaber_b200/librs_sched.so
  for rep in 1 2; do
    RS_SCHED_LIB=$lib timeout 300 python bench.py --kernel-only --steps 10 --warmup 3 2>>$O/bench_ko.err | sed "s/^/$v /" >> $O/bench_ko.jsonl
  done
  RS_SCHED_LIB=$lib timeout 300 python tools/sort_bench.py --check 2>&1 | sed "s/^/$v /" >> $O/sort_bench.txt
done
timeout 900 python bench.py > $O/bench_product.json 2> $O/bench_product.err; echo "bench rc=$?" >> $O/bench_product.err
tail -3 $O/pytest.log; cat $O/bench_ko.jsonl | cut -c1-200; cat $O/sort_bench.txt
