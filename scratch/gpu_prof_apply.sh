#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"gemm_stream|gemm_tma" -c 2 -o /tmp/prof_apply python scratch/prof_apply.py > gpurun_out/ncu_apply.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_apply.ncu-rep --page raw --csv > gpurun_out/prof_apply_raw.csv 2>/dev/null
ncu -i /tmp/prof_apply.ncu-rep --page source --csv > gpurun_out/prof_apply_src.csv 2>/dev/null
ncu -i /tmp/prof_apply.ncu-rep --page details > gpurun_out/prof_apply_details.txt 2>/dev/null
ls -la gpurun_out | tail -5
