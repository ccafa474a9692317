#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_n1.log 2>gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/bench_n2.log 2>gpurun_out/bench_n2.err; tail -1 gpurun_out/bench_n2.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'])"
tail -3 gpurun_out/bench_n2.err
