#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -12 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke 2>&1 | tail -3
python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench.log
python bench.py --steps 2 --warmup 3 --workload kitti --pairs 16 --pairs-per-launch 8 --distinct 8 --no-cpu-baseline > gpurun_out/bench_kitti.log 2>&1; echo "kitti rc=$?"; tail -1 gpurun_out/bench_kitti.log | cut -c1-700
