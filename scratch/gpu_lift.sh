#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q tests/test_e2pn_gpu.py -k "lifted or backbone" 2>&1 | tail -3
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kpconv_lift_kernel" -c 1 -o gpurun_out/prof_lift python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_lift.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_lift.ncu-rep --page details 2>/dev/null | grep -E "Duration|Executed Ipc|Issue Slots Busy|Registers Per|Achieved Occupancy|Theoretical Occupancy|L1/TEX Hit|L2 Hit|DRAM Throughput|No Eligible" | head -30
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_q.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_q.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('pairs/s', d['value'], 'e2e', d['e2e']['value'], {k: v for k, v in d['roofline']['per_entry_point_ms'].items() if v})"
