#!/bin/bash
mkdir -p gpurun_out
for cfg in "3 22" "4 16"; do
set -- $cfg
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-extra --streams $1 --pairs-per-launch $2 --pairs $(( $1 * $2 )) > gpurun_out/bench_n8s$1.log 2>gpurun_out/bench_n8s$1.err; tail -1 gpurun_out/bench_n8s$1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=8 streams $1 x $2:', d['value'], d['ms_per_step'], 'per-GPU pairs/s', d['value']/8)"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra --streams $1 --pairs-per-launch $2 --pairs $(( $1 * $2 )) > gpurun_out/bench_n1s$1.log 2>&1; tail -1 gpurun_out/bench_n1s$1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=1 streams $1 x $2:', d['value'], d['ms_per_step'])"
done
