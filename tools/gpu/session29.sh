#!/bin/bash
# headline A/B: working tree against the reference build (HEAD of the session start), kernel only, same box
O=gpurun_out/s29
mkdir -p $O
for lib in cur old cur old; do
  export RS_SCHED_LIB=$PWD/build/librs_$lib.so
  timeout 300 python bench.py --kernel-only --steps 12 --warmup 5 2>>$O/err | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('$lib', round(d['value']/1e6,3))"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sort or golden or headline or fixed" 2>&1 | tail -2
