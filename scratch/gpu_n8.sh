#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_n$N.log 2>gpurun_out/bench_n$N.err; tail -1 gpurun_out/bench_n$N.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=$N', d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
tail -2 gpurun_out/bench_n$N.err
