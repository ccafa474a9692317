#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kpconv_fused" -c 3 -o gpurun_out/prof_fused python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_full.log 2>&1; echo "ncu2 rc=$?"
ncu -i gpurun_out/prof_fused.ncu-rep --page raw --csv > gpurun_out/prof_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_fused.ncu-rep --page source --csv > gpurun_out/prof_fused_src.csv 2>/dev/null
ls -la gpurun_out
