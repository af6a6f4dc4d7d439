#!/bin/bash
# round 2, GPU session 14: scan-free Vogel line scans (tests, speed) + ncu capture of the id 103 kernel
O=gpurun_out/s14
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_two_bearers_gpu.py tests/test_dropin_gpu.py -m gpu -q -k "103 or fuzz or golden or reference_record or two_bearers or random" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --ids 103 2>>$O/sweep.err > $O/sweep_ids.jsonl
RS_NO_SPLIT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:rs_tti_kernel -s 8 -c 1 -o $O/r02_full_id103 python tools/sweep_bench.py --only ids --ids 103 --launches 1 > $O/ncu_103.log 2>&1
tail -3 $O/pytest.log; cat $O/sweep_ids.jsonl | cut -c1-200
