#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.log 2>&1; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f launches %d' % (d['value'], d['e2e']['value'], d['gpu_launches']))
print(d['roofline']['per_entry_point_ms'])
PY
