#!/bin/bash
timeout 700 python -m pytest tests/test_points_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.log 2>&1; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
pe = d['roofline']['per_entry_point_ms']
print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), 'radius', pe['se3et_radius_neighbors'], 'gn_double', pe['se3et_groupnorm_double'], 'gram', pe['se3et_linear_gnstats_gram'])
PY
