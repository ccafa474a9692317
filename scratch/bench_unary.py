"""Micro-benchmark of the Linear + GroupNorm passes at the backbone's shapes (32 stacked pairs)."""
import sys, torch
sys.path.insert(0, '.')
from se3et_b200.ops import gemm as G
from se3et_b200 import _lib
dev = torch.device('cuda')
P = 32
def seg_for(points):
    per = points // P
    return torch.arange(0, P + 1, dtype=torch.int64, device=dev) * per, per * P
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
shapes = [(917000, 32, 128), (917000, 64, 128), (266000, 32, 128), (266000, 64, 256), (266000, 128, 256), (74000, 128, 512), (74000, 256, 512)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for pts, k, n in shapes:
    seg, pts = seg_for(pts)
    m = pts * 6
    a = torch.randn(m, k, device=dev).to(torch.bfloat16)
    w = (torch.randn(n, k, device=dev) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(n, device=dev)
    gamma, beta = torch.randn(n, device=dev), torch.randn(n, device=dev)
    res = {}
    if k in (32, 64, 128):
        res['gram'] = timeit(lambda: G.linear_gn_stats_gram(a, w, b, 32, seg, 6))
    G._GRAM['on'] = False
    res['gnstats'] = timeit(lambda: G.linear_gn_stats(a, w, b, 32, seg, 6, store=False))
    G._GRAM['on'] = True
    _, st = G.linear_gn_stats(a, w, b, 32, seg, 6, store=False)
    L = _lib.lib()
    resid = torch.randn(m, n, device=dev).to(torch.bfloat16)
    for on in (0, 1):
        L.se3et_gemm_set_stream_apply(on)
        tag = 'stream' if on else 'tile'
        res['apply_' + tag] = timeit(lambda: G.linear_gn_apply(a, w, b, st, gamma, beta, 1e-5, 0.1, 32, seg, 6))
        res['applyres_' + tag] = timeit(lambda: G.linear_gn_apply(a, w, b, st, gamma, beta, 1e-5, 0.1, 32, seg, 6, resid=resid))
        if k <= 128:
            a2 = torch.randn(m, 2 * k, device=dev).to(torch.bfloat16)
            w2 = (torch.randn(n, 2 * k, device=dev) / k ** 0.5).to(torch.bfloat16)
            _, st2 = G.linear_gn_stats(a2, w2, b, 32, seg, 6, store=False)
            res['dual_' + tag] = timeit(lambda: G.linear_gn_apply_dual(a, w, b, st, gamma, beta, a2, w2, b, st2, gamma, beta, 1e-5, 0.1, 32, seg, 6))
    gb_in = m * k * 2 / 1e9
    gb_out = m * n * 2 / 1e9
    print('M=%8d K=%3d N=%4d  in %.2f GB out %.2f GB | ' % (m, k, n, gb_in, gb_out) +
          '  '.join('%s %.3f' % (kk, v) for kk, v in res.items()))
