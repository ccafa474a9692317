#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))
print({k: v for k, v in d['roofline']['per_entry_point_ms'].items() if v})
for k, v in d['extra_workloads'].items():
    print(k, {kk: vv for kk, vv in v.items() if kk != 'config'}, v.get('config', {}).get('neighbor_limits'))
PY
