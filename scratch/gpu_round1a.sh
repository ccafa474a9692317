#!/bin/bash
# tests + bench + ncu launch list + ncu full of the top kernels (one 16-pair launch sequence)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" 
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kpconv_gather|groupnorm|gemm_tma|geo_embed_project|flash" -c 120 -o gpurun_out/prof_r1a python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ls -la gpurun_out
