#!/bin/bash
timeout 600 python -m pytest tests/test_sinkhorn_gpu.py tests/test_partition_gpu.py -m gpu -x -q 2>&1 | tail -5
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -x -q 2>&1 | tail -5
python - <<'PY'
import torch, sys
sys.path.insert(0, '.')
from se3et_b200.modules.sinkhorn import log_optimal_transport
dev = torch.device('cuda')
s = torch.randn(8192, 64, 64, device=dev) * 2
a = torch.tensor(1.0, device=dev)
for _ in range(2): log_optimal_transport(s, a, 100)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): log_optimal_transport(s, a, 100)
e1.record(); torch.cuda.synchronize()
print('log-sinkhorn 8192 x 65 x 65, 100 iterations: %.3f ms' % (e0.elapsed_time(e1) / 5))
PY
