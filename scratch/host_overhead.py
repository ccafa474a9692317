"""Host time of one launch sequence: forward_stacked on two tiny pairs (GPU work negligible), wall clock per call."""
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
from se3et_b200 import synthetic
from se3et_b200.model import make_cfg, create_model
dev = torch.device('cuda:0')
cfg = make_cfg('se3eti.3dmatch')
torch.manual_seed(0)
model = create_model(cfg).to(dev).eval()
pairs = [synthetic.make_3dmatch_pair(i, target_points=3000) for i in range(2)]
lens = np.array([len(c) for p in pairs for c in (p['ref_points'], p['src_points'])], dtype=np.int64)
pts = torch.from_numpy(np.concatenate([c for p in pairs for c in (p['ref_points'], p['src_points'])])).to(dev)
lens_t = torch.from_numpy(lens)
for _ in range(3):
    model.forward_stacked(pts, lens_t)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    model.forward_stacked(pts, lens_t)
torch.cuda.synchronize()
print('host ms per launch sequence (2 tiny pairs): %.2f' % ((time.perf_counter() - t0) / 10 * 1e3))
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for _ in range(5):
    model.forward_stacked(pts, lens_t)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(35)
