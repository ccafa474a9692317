import sys, torch
sys.path.insert(0, '.')
from se3et_b200.ops import gemm as G
from se3et_b200 import _lib
dev = torch.device('cuda')
P = 32; pts = 917000 // P * P; m = pts * 6; k = 32; n = 128
seg = torch.arange(0, P + 1, dtype=torch.int64, device=dev) * (pts // P)
a = torch.randn(m, k, device=dev).to(torch.bfloat16)
w = (torch.randn(n, k, device=dev) / k ** 0.5).to(torch.bfloat16)
b = torch.randn(n, device=dev); gamma = torch.randn(n, device=dev); beta = torch.randn(n, device=dev)
_, st = G.linear_gn_stats(a, w, b, 32, seg, 6, store=False)
L = _lib.lib()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for on in (0, 1):
    L.se3et_gemm_set_stream_apply(on)
    G.linear_gn_apply(a, w, b, st, gamma, beta, 1e-5, 0.1, 32, seg, 6)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
