#!/bin/bash
# round 2, GPU session 16: sort with scan units (any warp scans a unit of a long range): tests, A/B against the previous build
O=gpurun_out/s16
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sort or golden or reference_record or headline or sweep or fixed" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for lib in new old; do
  if [ $lib = old ]; then export RS_SCHED_LIB=$PWD/build/librs_old.so; fi
  timeout 300 python tools/sort_bench.py > $O/sort_$lib.txt 2>&1
  timeout 300 python bench.py --kernel-only --steps 10 --warmup 5 > $O/bench_$lib.json 2>$O/bench_$lib.err
  RS_NO_SPLIT=1 timeout 300 python bench.py --kernel-only --steps 10 --warmup 5 > $O/bench_nosplit_$lib.json 2>>$O/bench_$lib.err
  echo "== $lib"; cat $O/sort_$lib.txt; cut -c1-120 $O/bench_$lib.json $O/bench_nosplit_$lib.json
done
