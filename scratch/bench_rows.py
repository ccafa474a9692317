"""Micro-benchmark: se3et_kpconv_rows vs se3et_kpconv_fused on the levels of P stacked synthetic 3DMatch-shaped pairs."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from se3et_b200 import synthetic
from se3et_b200.precompute import precompute_data_stack_mode
from se3et_b200.ops import e2pn_ops as K
from se3et_b200.modules import e2pn as M
dev = torch.device('cuda')
P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
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
for lvl, cin, strided in ((0, 32, False), (0, 32, True), (1, 64, False), (1, 64, True), (2, 128, False), (3, 256, False)):
    s = dd['points'][lvl]
    q = dd['points'][lvl + 1] if strided else s
    nb = (dd['subsampling'][lvl] if strided else dd['neighbors'][lvl]).contiguous()
    conv = M.KPConvInterSO3(15, 6, cin, cin, 0.05 * 2 ** lvl, 0.0625 * 2 ** lvl, non_sep_conv=True, rot_by_permute=True, quotient_factor=4).to(dev)
    x = torch.randn(s.shape[0], 6, cin, device=dev).to(torch.bfloat16)
    w = conv._w_fused(); wr = conv._w_rows()
    yf = K.kpconv_fused(q, s, nb, x, w, conv.kernel_points, conv.KP_extent)[0]
    yr = K.kpconv_rows(q, s, nb, x, wr, conv.kernel_points, conv.KP_extent)
    err = ((yf - yr).norm() / yf.norm()).item()
    t0 = timeit(lambda: K.kpconv_fused(q, s, nb, x, w, conv.kernel_points, conv.KP_extent))
    t1 = timeit(lambda: K.kpconv_rows(q, s, nb, x, wr, conv.kernel_points, conv.KP_extent))
    fl = q.shape[0] * 432 * cin * cin * 2
    print('level %d%s nq %d H %d cin %d: fused %.3f ms (%.0f TF/s)  rows %.3f ms (%.0f TF/s)  rel diff %.2e' % (
        lvl, 's' if strided else ' ', q.shape[0], nb.shape[1], cin, t0, fl / t0 * 1e-9, t1, fl / t1 * 1e-9, err), flush=True)
