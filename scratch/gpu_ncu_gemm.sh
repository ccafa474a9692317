#!/bin/bash
mkdir -p gpurun_out
python scratch/prof_gemm.py
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_tma" -c 2 -o gpurun_out/prof_gemm python scratch/prof_gemm.py > gpurun_out/ncu_gemm.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_gemm.ncu-rep --page raw --csv > gpurun_out/prof_gemm_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_gemm.ncu-rep --page source --csv > gpurun_out/prof_gemm_src.csv 2>/dev/null
rm -f gpurun_out/prof_gemm.ncu-rep
ls -la gpurun_out | head -20
