#!/bin/bash
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python scratch/bench_unary.py 2>&1 | cut -c1-260 | head -3
for mode in lift_off lift_on; do
python - $mode <<'PY'
import sys, json, subprocess
mode = sys.argv[1]
code = """
import sys
sys.argv = ['bench.py', '--steps', '5', '--warmup', '3']
from se3et_b200.modules import e2pn
if %r == 'lift_off': e2pn._GFLAGS['lifted_kernel'] = False
import runpy
runpy.run_path('bench.py', run_name='__main__')
""" % mode
out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    pe = d['roofline']['per_entry_point_ms']
    print(mode, 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), pe)
except Exception as e:
    print(mode, 'failed', e, out.stdout[-500:], out.stderr[-1500:])
PY
done
