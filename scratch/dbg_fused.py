import sys, ctypes
sys.path.insert(0, '/root/repo')
import torch
from se3et_b200 import _lib
from se3et_b200.ops import e2pn_ops as K
L = _lib.lib()
for bn in (16, 32, 64, 128):
    o = (ctypes.c_int * 5)()
    rc = L.se3et_kpconv_fused_attrs(bn, o)
    print(bn, rc, list(o), L.se3et_last_error())
dev = 'cuda:0'
torch.manual_seed(0)
nq, h, cin, cout = 100, 38, 16, 32
q = torch.rand(nq, 3, device=dev) * 0.2
nb = torch.randint(0, nq + 1, (nq, h), device=dev)
x = torch.randn(nq, 6, cin, device=dev).bfloat16()
w = torch.randn(cout, 36 * cin, device=dev).bfloat16()
kp = torch.rand(15, 3, device=dev) * 0.04
try:
    y, _ = K.kpconv_fused(q, q, nb, x, w, kp, 0.05)
    torch.cuda.synchronize()
    print('ok', y.abs().mean().item())
except Exception as e:
    print('ERR', e)
