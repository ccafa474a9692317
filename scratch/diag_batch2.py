import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import helpers
from test_e2pn_gpu import _build_backbone, rel_err
from se3et_b200.precompute import precompute_data_stack_mode
DEV='cuda:0'
S = helpers.SMALL_CFG
net,_ = _build_backbone()
g = np.load('/root/repo/tests/golden/model_small.npz')
pts_a, lens_a = g['in_points'], g['in_lengths']
pts_b, lens_b = helpers.small_pair(index=12, crop=1.1)
def run(pts, lens):
    d = precompute_data_stack_mode(torch.from_numpy(pts).to(DEV), torch.from_numpy(lens).to(DEV), 4, S['init_voxel'], S['init_radius'], [38,36,36,38])
    feats = {}
    hooks = []
    for name, m in net.named_modules():
        if name.count('.')<=1 and name:
            hooks.append(m.register_forward_hook(lambda mod, inp, out, name=name: feats.__setitem__(name, out.float() if torch.is_tensor(out) else None)))
    with torch.no_grad():
        out = net(torch.ones(len(pts),1,device=DEV), d)
    for h in hooks: h.remove()
    return d, feats
dB, fB = run(pts_b, lens_b)
d2, f2 = run(np.concatenate([pts_a, pts_b]), np.concatenate([lens_a, lens_b]))
print([w.tolist() for w in d2['subsampling_width']], [s.shape for s in dB['subsampling']])
for k in fB:
    if fB[k] is None: continue
    n = fB[k].shape[0]
    print('batched vs single B', k, n, rel_err(f2[k][-n:], fB[k]))
