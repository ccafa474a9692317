#!/bin/bash
python bench.py --steps 3 --warmup 3 --workload kitti --pairs 16 --pairs-per-launch 8 --no-cpu-baseline > gpurun_out/bench_kitti.log 2>gpurun_out/bench_kitti.err; tail -3 gpurun_out/bench_kitti.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_kitti.log').read().strip().splitlines()[-1])
print('kitti value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), d['config'].get('neighbor_limits'))
print({k: v for k, v in d['roofline']['per_entry_point_ms'].items() if v})
PY
