#!/bin/bash
# round 2, GPU session 5: direct-metric path for big slices (tests + sweep A/B)
O=gpurun_out/s5
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
PTS="5,40;10,10;10,20;20,10;20,20;20,40;30,10;50,10;50,40;40,20;15,15"
timeout 900 python tools/sweep_bench.py --only sweep --points "$PTS" 2>>$O/sweep.err | sed "s/^/direct /" >> $O/sweep_direct_ab.jsonl
RS_NO_DIRECT=1 timeout 900 python tools/sweep_bench.py --only sweep --points "$PTS" 2>>$O/sweep.err | sed "s/^/table /" >> $O/sweep_direct_ab.jsonl
tail -3 $O/pytest.log; python - <<'PY'
import json
for l in open('gpurun_out/s5/sweep_direct_ab.jsonl'):
    v,j=l.split(' ',1); d=json.loads(j); print(v, d['label'], round(d['cell_ttis_per_s']/1e6,3), d['cells'], d['smem_bytes_per_cta'])
PY
