#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_e2pn_gpu.py -m gpu -x -q -k "gram or dual" 2>&1 | tail -8
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_e2pn_gpu.py -m gpu -x -q 2>&1 | tail -3
for mode in gram_on; do
python - $mode <<'PY'
import sys, json, subprocess
mode = sys.argv[1]
code = """
import sys
sys.argv = ['bench.py', '--steps', '5', '--warmup', '3']
from se3et_b200.ops import gemm
if %r == 'gram_off': gemm._GRAM['on'] = False
import runpy
runpy.run_path('bench.py', run_name='__main__')
""" % mode
out = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    pe = d['roofline']['per_entry_point_ms']
    print(mode, 'value %.1f e2e %.1f' % (d['value'], d['e2e']['value']), {k: v for k, v in pe.items() if 'gemm' in k or 'gram' in k})
    print('   frac', d['roofline']['per_entry_point_frac'])
except Exception as e:
    print(mode, 'failed', e, out.stdout[-500:], out.stderr[-1500:])
PY
done
