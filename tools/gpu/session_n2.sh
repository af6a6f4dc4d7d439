#!/bin/bash
# round 2, 2-GPU session: rs_batch --gpus 2 == --gpus 1, bench at N=2 (weak) and BASELINE configs[4] (--cells-total 65536)
O=gpurun_out/n2_final
mkdir -p $O
nvidia-smi -L > $O/gpus.txt
timeout 900 python -m pytest tests/test_host_api_gpu.py -m gpu -q -k "two_gpus or reduce" > $O/pytest.log 2>&1; echo "rc=$?" >> $O/pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > $O/bench_n2.json 2> $O/bench_n2.err; echo "rc=$?" >> $O/bench_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --cells-total 65536 --steps 18 --warmup 3 > $O/bench_n2_cells65536.json 2> $O/bench_n2_cells65536.err; echo "rc=$?" >> $O/bench_n2_cells65536.err
timeout 600 ./radiosaber_b200/rs_batch --algo 9 --config tests/data/cfg20x5.json --cells 65536 --ttis 1000 --gpus 2 > $O/rs_batch_65536x1000_2gpu.json 2> $O/rs_batch.err
tail -3 $O/pytest.log; cut -c1-300 $O/bench_n2.json; tail -2 $O/bench_n2.err; cut -c1-300 $O/bench_n2_cells65536.json; tail -2 $O/bench_n2_cells65536.err; cut -c1-200 $O/rs_batch_65536x1000_2gpu.json; tail -2 $O/rs_batch.err
