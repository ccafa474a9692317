#!/bin/bash
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2>gpurun_out/bench_n2.err; tail -5 gpurun_out/bench_n2.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_n2.log').read().strip().splitlines()[-1])
print('N=2 value %.1f e2e %.1f ms/step %.1f' % (d['value'], d['e2e']['value'], d['ms_per_step']))
for k, v in d['extra_workloads'].items():
    print(k, {kk: vv for kk, vv in v.items() if kk not in ('config', 'top_entry_points_ms')})
PY
