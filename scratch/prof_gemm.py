import sys
sys.path.insert(0, '/root/repo')
import torch
from se3et_b200.ops.gemm import linear_gn_stats, linear_gn_apply
dev = 'cuda:0'
torch.manual_seed(0)
M, K, N, G = 3362000, 32, 128, 32
a = torch.randn(M, K, device=dev).bfloat16()
w = torch.randn(N, K, device=dev).bfloat16()
bias = torch.randn(N, device=dev)
gamma = torch.randn(N, device=dev); beta = torch.randn(N, device=dev)
resid = torch.randn(M, N, device=dev).bfloat16()
pts = M // 6 // 16
seg = torch.arange(0, 17, device=dev, dtype=torch.int64) * pts
seg[-1] = M // 6
for i in range(2):
    if i == 1:
        torch.cuda.profiler.start()
    _, st = linear_gn_stats(a, w, bias, G, seg, 6, store=False)
    out = linear_gn_apply(a, w, bias, st, gamma, beta, 1e-5, 0.1, G, seg, 6, resid=resid)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
e0.record(); _, st = linear_gn_stats(a, w, bias, G, seg, 6, store=False); e1.record()
out = linear_gn_apply(a, w, bias, st, gamma, beta, 1e-5, 0.1, G, seg, 6, resid=resid); e2.record()
torch.cuda.synchronize()
print('stats %.1f us, apply %.1f us' % (e0.elapsed_time(e1) * 1e3, e1.elapsed_time(e2) * 1e3))
