"""One launch sequence of the hot path for ncu: `python scratch/profile_step.py [variant] [pairs] [passes]`.
Pass 0 is the warm-up (skip its launches with ncu -s), the later passes are what gets profiled."""
import sys
sys.path.insert(0, '/root/repo')
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from se3et_b200 import synthetic, _lib
from se3et_b200.model import make_cfg, create_model

dev = torch.device('cuda:0')
variant = sys.argv[1] if len(sys.argv) > 1 else 'se3eti.3dmatch'
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 16
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = make_cfg(variant)
torch.manual_seed(0)
model = create_model(cfg).to(dev).eval()
pairs = [synthetic.make_3dmatch_pair(i) for i in range(npairs)]
lens = np.array([len(c) for p in pairs for c in (p['ref_points'], p['src_points'])], dtype=np.int64)
pts = torch.from_numpy(np.concatenate([c for p in pairs for c in (p['ref_points'], p['src_points'])])).to(dev)
L = _lib.lib()
if os.environ.get('SE3ET_RADIUS_MODE'):
    L.se3et_radius_set_mode(int(os.environ['SE3ET_RADIUS_MODE']))
for i in range(passes):
    if i == 1:
        torch.cuda.profiler.start()
    L.enabled = True
    L.reset()
    model.forward_stacked(pts, torch.from_numpy(lens))
    torch.cuda.synchronize()
    L.enabled = False
    print('pass', i, 'launches', L.launches(), flush=True)
torch.cuda.profiler.stop()
