#!/bin/bash
# round 2, GPU session 15: FixedShape instantiations for ids 10 / 101 / 103 (tests: fixed == general == oracle), A/B sweep
O=gpurun_out/s15
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_fuzz_gpu.py tests/test_two_bearers_gpu.py -m gpu -q -k "10 or 101 or 103 or fixed or fuzz or golden" > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
timeout 600 python tools/sweep_bench.py --only ids --ids 10,101,103 2>>$O/sweep.err > $O/sweep_fixed.jsonl
RS_NO_FIXED_SHAPE=1 timeout 600 python tools/sweep_bench.py --only ids --ids 10,101,103 2>>$O/sweep.err > $O/sweep_general.jsonl
tail -3 $O/pytest.log; cut -c1-60,330-380 $O/sweep_fixed.jsonl $O/sweep_general.jsonl
