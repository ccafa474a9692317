#!/bin/bash
timeout 1500 python -m pytest tests/test_e2pn_gpu.py tests/test_gemm_gpu.py tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 5 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/bench_quick.log 2>gpurun_out/bench_quick.err; tail -3 gpurun_out/bench_quick.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_quick.log').read().strip().splitlines()[-1])
print('value %.1f e2e %.1f' % (d['value'], d['e2e']['value']))
print({k: v for k, v in d['roofline']['per_entry_point_ms'].items() if v})
PY
python scratch/bench_train.py 2>&1 | grep -E "recompute|Error|error" | tail -4
