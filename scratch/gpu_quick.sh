#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q tests/test_e2pn_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('pairs/s', d['value'], 'e2e', d['e2e']['value'], d['roofline']['per_entry_point_ms'])"
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
