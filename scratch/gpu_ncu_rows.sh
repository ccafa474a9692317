#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kpconv_rows -c 2 -o gpurun_out/prof_rows -f python scratch/bench_rows.py 8 > gpurun_out/ncu_rows.log 2>&1
tail -3 gpurun_out/ncu_rows.log
