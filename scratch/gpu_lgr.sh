#!/bin/bash
timeout 900 python -m pytest tests/test_registration_gpu.py -m gpu -x -q 2>&1 | tail -15
python - <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from se3et_b200 import synthetic
from se3et_b200.calibrate import calibrate_neighbors_stack_mode
pairs = [synthetic.make_kitti_pair(i) for i in range(3)]
cl = [(p['ref_points'], p['src_points']) for p in pairs]
print('kitti calibrated limits', calibrate_neighbors_stack_mode(cl, 5, 0.3, 4.25 * 0.3))
pairs = [synthetic.make_3dmatch_pair(i) for i in range(3)]
cl = [(p['ref_points'], p['src_points']) for p in pairs]
print('3dmatch calibrated limits', calibrate_neighbors_stack_mode(cl, 4, 0.025, 0.0625))
PY
