#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_q.csv python scratch/profile_step.py se3eti.3dmatch 32 2 > gpurun_out/ncu_launch_q.log 2>&1; echo "ncu1 rc=$?"
