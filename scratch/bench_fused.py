"""Micro-benchmark of se3et_kpconv_fused with / without the GroupNorm statistics epilogue (32 stacked pairs, level 0)."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from se3et_b200 import synthetic
from se3et_b200.precompute import precompute_data_stack_mode
from se3et_b200.ops import e2pn_ops as K
from se3et_b200.modules import e2pn as M
dev = torch.device('cuda')
P = 16
pairs = [synthetic.make_3dmatch_pair(1000 + i) for i in range(P)]
pts = np.concatenate([np.concatenate([p['ref_points'], p['src_points']]) for p in pairs]).astype(np.float32)
lens = np.array([n for p in pairs for n in (len(p['ref_points']), len(p['src_points']))], np.int64)
dd = precompute_data_stack_mode(torch.from_numpy(pts).to(dev), torch.from_numpy(lens).to(dev), 4, 0.025, 0.0625, [38, 36, 36, 38], backbone_only=True)
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for lvl, cin in ((0, 32), (1, 64), (2, 128)):
    q = dd['points'][lvl]; nb = dd['neighbors'][lvl].contiguous()
    conv = M.KPConvInterSO3(15, 6, cin, cin, 0.05 * 2 ** lvl, 0.0625 * 2 ** lvl, non_sep_conv=True, rot_by_permute=True, quotient_factor=4).to(dev)
    x = torch.randn(q.shape[0], 6, cin, device=dev).to(torch.bfloat16)
    nq = q.shape[0]
    seg = torch.tensor([0, nq // 2, nq], dtype=torch.int64, device=dev)
    w = conv._w_fused()
    t0 = timeit(lambda: K.kpconv_fused(q, q, nb, x, w, conv.kernel_points, conv.KP_extent))
    t1 = timeit(lambda: K.kpconv_fused(q, q, nb, x, w, conv.kernel_points, conv.KP_extent, gn=(32, seg)))
    print('level %d nq %d cin %d: no stats %.3f ms, with stats %.3f ms' % (lvl, nq, cin, t0, t1))
