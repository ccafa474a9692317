"""Micro-benchmark of the Linear+GroupNorm statistics passes on the backbone's shapes (32 stacked pairs)."""
import sys, torch, numpy as np
sys.path.insert(0, '.')
from se3et_b200.ops import gemm as G
dev = torch.device('cuda')
def timeit(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
P = 32
pts = {0: 28663, 1: 8527, 2: 2388, 3: 680}
shapes = [(0, 64, 32), (0, 64, 128), (0, 128, 32), (1, 128, 64), (1, 128, 256), (1, 256, 64), (1, 256, 64), (2, 256, 128), (2, 256, 512), (2, 512, 128), (2, 512, 128), (3, 512, 256), (3, 512, 1024), (3, 1024, 256),
          (0, 32, 128), (1, 32, 128), (1, 64, 256), (2, 64, 256), (2, 128, 512), (3, 128, 512), (3, 256, 1024)]
tot = {'stream': 0, 'old': 0}
for lvl, k, n in shapes:
    rows = 6 * pts[lvl] * P
    seg = torch.arange(P + 1, device=dev, dtype=torch.int64) * pts[lvl]
    a = torch.randn(rows, k, device=dev).to(torch.bfloat16)
    w = (torch.randn(n, k, device=dev) / k ** 0.5).to(torch.bfloat16)
    b = torch.randn(n, device=dev)
    t_s = timeit(lambda: G.linear_gn_stats_stream(a, w, b, 32, seg, 6))
    G.STATS_MODE['stream'] = False
    t_o = timeit(lambda: G.linear_gn_stats(a, w, b, 32, seg, 6, store=False))
    G.STATS_MODE['stream'] = True
    gb = rows * k * 2 / 1e9
    tot['stream'] += t_s; tot['old'] += t_o
    print('lvl %d rows %8d k %4d n %4d: stream %.3f ms (%.0f GB/s)  old %.3f ms (%.0f GB/s)' % (lvl, rows, k, n, t_s, gb / t_s * 1e3, t_o, gb / t_o * 1e3), flush=True)
print(tot)
