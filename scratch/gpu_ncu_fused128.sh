#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"kpconv_fused" --launch-skip 5 --launch-count 2 -o /tmp/prof_f128 python scratch/profile_step.py se3eti.3dmatch 16 2 > gpurun_out/ncu_f128.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_f128.ncu-rep --page source --csv > gpurun_out/prof_f128_src.csv 2>/dev/null
ncu -i /tmp/prof_f128.ncu-rep --page details > gpurun_out/prof_f128_details.txt 2>/dev/null
python scratch/src_ops.py gpurun_out/prof_f128_src.csv 0 > gpurun_out/prof_f128_ops.txt 2>&1
python scratch/src_hist.py gpurun_out/prof_f128_src.csv 0 > gpurun_out/prof_f128_hist.txt 2>&1
rm -f gpurun_out/prof_f128_src.csv
ls -la gpurun_out | tail -6
