#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_model_gpu.py -m gpu -x -q -s -k "headline or kitti_30k or 32_stacked" 2>&1 | grep -E "PARITY|passed|failed|Error|error" | head -20
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.log 2>gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))
print({k: v for k, v in d['roofline']['per_entry_point_ms'].items() if v})
print(d['roofline']['per_entry_point_frac'])
print(json.dumps(d['extra_workloads'], indent=1)[:3000])
PY
