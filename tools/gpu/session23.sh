#!/bin/bash
O=gpurun_out/s23
mkdir -p $O
RS_NO_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_tti_kernel -s 8 -c 1 -o $O/r02_full_id103_fixed python tools/sweep_bench.py --only ids --ids 103 --launches 1 > $O/ncu_103.log 2>&1
tail -2 $O/ncu_103.log
