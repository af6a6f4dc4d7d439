#!/bin/bash
# round 2, final build on 8 GPUs: the weak-scaling bench line only (the first session's run has configs[4] and rs_batch)
O=gpurun_out/n8_final
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc=$?" >> $O/bench_n8.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/n8_final/bench_n8.json').readline())
print('n8', round(d['value']/1e6,2), 'e2e', round(d['e2e']['value']/1e6,2), {k: round(v['value']/1e6,2) for k,v in d['e2e']['variants'].items()})
PY
tail -2 $O/bench_n8.err
