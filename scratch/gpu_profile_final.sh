#!/bin/bash
# round-end evidence: tests, bench, launch list of one 32-pair launch sequence, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_launch.log 2>&1; echo "ncu1 rc=$?"
ncu --set full --clock-control none --profile-from-start off -k regex:"kpconv_fused|kpconv_cin1|gemm_tma|gemm_stream|gemm_dual|gram_kernel|geo_embed_lookup|geo_embed_project|radius_query|kpconv_gather|flash_attention|groupnorm_double|maxpool" -c 110 -o /tmp/prof_all python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/prof_all.ncu-rep
