import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from se3et_b200 import synthetic
from se3et_b200.model import make_cfg, create_model
dev = torch.device('cuda:0')
cfg = make_cfg('se3eti.3dmatch'); torch.manual_seed(0)
model = create_model(cfg).to(dev).eval()
pairs = [synthetic.make_3dmatch_pair(i) for i in range(16)]
def mk(group):
    lens = np.array([len(c) for p in group for c in (p['ref_points'], p['src_points'])], dtype=np.int64)
    pts = torch.from_numpy(np.concatenate([c for p in group for c in (p['ref_points'], p['src_points'])])).to(dev)
    return pts, torch.from_numpy(lens)
for ppl in (16, 8):
    groups = [pairs[i:i + ppl] for i in range(0, 16, ppl)] * (64 // 16)
    inputs = [mk(g) for g in groups]
    for ns in (1, 2, 3, 4):
        for _ in range(2):
            model.forward_stacked_concurrent(inputs, num_streams=ns)
        torch.cuda.synchronize(); t = time.perf_counter()
        for _ in range(3):
            model.forward_stacked_concurrent(inputs, num_streams=ns)
        torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
        print('pairs/launch %d streams %d: %.1f ms per 64 pairs -> %.1f pairs/s' % (ppl, ns, dt * 1e3, 64 / dt), flush=True)
