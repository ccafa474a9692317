#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q tests/test_points_gpu.py tests/test_model_gpu.py 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_q.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_q.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('pairs/s', d['value'], 'e2e', d['e2e']['value'])"
