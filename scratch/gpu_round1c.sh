#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kpconv_fused" -c 4 -o gpurun_out/prof_fused python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ncu -i gpurun_out/prof_fused.ncu-rep --page raw --csv > gpurun_out/prof_fused_raw.csv 2>/dev/null
ls -la gpurun_out
