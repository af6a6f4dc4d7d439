#!/bin/bash
# metric-table chunks of 40 UEs (8 slices) on the headline cell: tests, A/B against the session-start build (ratio was 1.008-1.009 before)
O=gpurun_out/s36
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py -m gpu -q -x > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
for lib in new old new old; do
  [ $lib = old ] && export RS_SCHED_LIB=$PWD/build/librs_old.so || unset RS_SCHED_LIB
  timeout 300 python bench.py --kernel-only --steps 12 --warmup 5 2>>$O/err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$lib', round(d['value']/1e6,3))"
done
unset RS_SCHED_LIB
timeout 600 python tools/sweep_bench.py --only sweep --points "50,2;20,2;30,5;40,5;15,5;10,5" 2>>$O/err | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print(d['label'], round(d['cell_ttis_per_s']/1e6,2))"
