#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"groupnorm_double_kernel" -c 8 -o gpurun_out/prof_gnd python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_gnd.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_gnd.ncu-rep --page details 2>/dev/null | grep -E "groupnorm_double_kernel|Duration|Executed Ipc Active|Issue Slots Busy|Registers Per|Achieved Occupancy|L2 Hit|DRAM Throughput|No Eligible|Memory Throughput" | head -80
