#!/bin/bash
# final evidence refresh: tests, smoke, bench, launch list, ncu --set full of the kernels added after the big capture
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
python bench.py > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --profile-from-start off -k regex:"linear_add_layernorm|flash_attention|gemm_tma" -c 40 -o /tmp/prof_ln python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_ln.log 2>&1; echo "ncu2 rc=$?"
ncu -i /tmp/prof_ln.ncu-rep --page raw --csv > gpurun_out/prof_ln_raw.csv 2>/dev/null
