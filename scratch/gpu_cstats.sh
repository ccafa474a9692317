#!/bin/bash
timeout 700 python -m pytest tests/test_e2pn_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -3
for mode in 0 1; do
python - $mode <<'PY'
import sys, json, subprocess
mode = int(sys.argv[1])
code = """
import sys
sys.argv = ['bench.py', '--steps', '5', '--warmup', '3', '--no-cpu-baseline']
from se3et_b200.modules import e2pn
e2pn._GFLAGS['conv_stats_stream'] = bool(%d)
import runpy
runpy.run_path('bench.py', run_name='__main__')
""" % mode
out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    pe = d['roofline']['per_entry_point_ms']
    print('conv_stats_stream', mode, 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), 'fused', pe['se3et_kpconv_fused'], 'gn_double', pe['se3et_groupnorm_double'])
except Exception as e:
    print(mode, 'failed', e, out.stdout[-500:], out.stderr[-1500:])
PY
done
