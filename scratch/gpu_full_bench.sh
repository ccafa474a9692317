#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.log 2>gpurun_out/bench_full.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_full.log | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('pairs/s', d['value'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['frac'])
for k, v in d['extra_workloads'].items(): print(k, v.get('value'), v.get('e2e'), v.get('top_entry_points_ms'))
print(d['cpu_baseline'])"
