#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_e2pn_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -4
python scratch/bench_unary.py
for mode in stream; do
python - $mode <<'PY'
import sys, json, subprocess
mode = sys.argv[1]
code = """
import sys
sys.argv = ['bench.py', '--steps', '5', '--warmup', '3']
from se3et_b200 import _lib
_lib.lib().se3et_gemm_set_stream_apply(%d)
import runpy
runpy.run_path('bench.py', run_name='__main__')
""" % (1 if mode == 'stream' else 0)
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
