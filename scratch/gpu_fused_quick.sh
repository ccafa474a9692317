#!/bin/bash
timeout 600 python -m pytest tests/test_e2pn_gpu.py -m gpu -x -q -k "fused or kpconv or equivariance" 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.log 2>&1; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), 'fused', d['roofline']['per_entry_point_ms']['se3et_kpconv_fused'])
PY
