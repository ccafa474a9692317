#!/bin/bash
for cfg in "2 32" "4 16" "3 22" "2 16" "1 64"; do
  set -- $cfg
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline --streams $1 --pairs-per-launch $2 > gpurun_out/bench_s.log 2>&1
  python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.loads(open('gpurun_out/bench_s.log').read().strip().splitlines()[-1])
    print('streams', sys.argv[1], 'ppl', sys.argv[2], 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))
except Exception as e:
    print('failed', sys.argv[1:], e, open('gpurun_out/bench_s.log').read()[-600:])
PY
done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload kitti 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('kitti', d['value'], d['e2e']['value'])"
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload 3dmatch-e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('3dmatch-e', d['value'], d['e2e']['value'])"
