#!/bin/bash
# round 2, 8-GPU session: bench at N=8 (weak, 4096 cells per GPU), BASELINE configs[4] as written (65536 cells x 1000+ TTIs over 8
# GPUs, final NCCL reduce), the all-C++ runner on the same configuration
O=gpurun_out/n8
mkdir -p $O
nvidia-smi -L > $O/gpus.txt; nproc >> $O/gpus.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?" >> $O/bench_n8.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --cells-total 65536 --steps 18 --warmup 3 > $O/bench_n8_cells65536.json 2> $O/bench_n8_cells65536.err; echo "rc=$?" >> $O/bench_n8_cells65536.err
timeout 600 ./radiosaber_b200/rs_batch --algo 9 --config tests/data/cfg20x5.json --cells 65536 --ttis 1000 --gpus 8 > $O/rs_batch_65536x1000_8gpu.json 2> $O/rs_batch.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --no-cpu-baseline > $O/bench_n4.json 2> $O/bench_n4.err; echo "rc=$?" >> $O/bench_n4.err
python - <<'PY'
import json
for f in ('bench_n8.json','bench_n8_cells65536.json','bench_n4.json'):
    try:
        d=json.loads(open('gpurun_out/n8/'+f).read().strip().splitlines()[-1])
        v=d['e2e']['variants']['packed_cqi_every_tti']
        print(f, round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), {k:round(x['value']/1e6,2) for k,x in d['e2e']['variants'].items()}, 'h2d', round(v['h2d_gbs_per_gpu'],1), 'ceil', round(v['h2d_ceiling_gbs_per_gpu'],1), d['parity_spot']['mismatches'], d['stats_reduce'][:20])
    except Exception as e:
        print(f, 'ERR', e)
print(open('gpurun_out/n8/rs_batch_65536x1000_8gpu.json').read()[:260])
PY
tail -2 $O/bench_n8.err $O/rs_batch.err
