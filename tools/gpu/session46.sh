#!/bin/bash
# round 2, last check of the final tree: full GPU suite, smoke, the bench line
O=gpurun_out/s46
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/pytest.log 2>&1; echo "pytest rc=$?" >> $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.log 2>&1; tail -1 $O/smoke.log
timeout 900 python bench.py > $O/bench.json 2>$O/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s46/bench.json').readline())
print('bench', round(d['value']/1e6,3), 'e2e', round(d['e2e']['value']/1e6,3), {k: round(v['value']/1e6,2) for k,v in d['e2e']['variants'].items()}, 'parity', d['parity_spot']['mismatches'], 'frac', round(d['roofline']['frac'],4), 'launches', d['gpu_launches'])
PY
