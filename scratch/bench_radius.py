import sys, torch, numpy as np
sys.path.insert(0, '.')
from se3et_b200 import synthetic, ext, _lib
from se3et_b200.ops import grid_subsample
dev = torch.device('cuda')
P = 32
pairs = [synthetic.make_3dmatch_pair(1000 + i) for i in range(P)]
pts = torch.from_numpy(np.concatenate([np.concatenate([p['ref_points'], p['src_points']]) for p in pairs]).astype(np.float32)).to(dev)
lens = torch.tensor([n for p in pairs for n in (len(p['ref_points']), len(p['src_points']))], dtype=torch.int64, device=dev)
p1, l1, _ = grid_subsample(pts, lens, torch.zeros_like(pts), voxel_size=0.05)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for name, q, s, ql, sl, r, w in (('L0 self', pts, pts, lens, lens, 0.0625, 38), ('L0->L1', p1, pts, l1, lens, 0.0625, 38), ('L1 self', p1, p1, l1, l1, 0.125, 36)):
    t = timeit(lambda: ext.radius_neighbors_raw(q, s, ql, sl, r, w))
    out, _, status = ext.radius_neighbors_raw(q, s, ql, sl, r, w)
    print('%s nq %d ns %d: %.3f ms  status %s' % (name, q.shape[0], s.shape[0], t, status[:8].tolist()), flush=True)
