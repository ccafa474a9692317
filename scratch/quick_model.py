import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from se3et_b200 import synthetic, _lib
from se3et_b200.model import make_cfg, create_model
dev = torch.device('cuda:0')
cfg = make_cfg(sys.argv[1] if len(sys.argv) > 1 else 'se3eti.3dmatch')
torch.manual_seed(0)
model = create_model(cfg).to(dev).eval()
print('params', sum(p.numel() for p in model.parameters() if p.requires_grad))
npairs = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pairs = [synthetic.make_3dmatch_pair(i) for i in range(npairs)]
clouds = [(p['ref_points'], p['src_points']) for p in pairs]
for rep in range(3):
    torch.cuda.synchronize(); t = time.time()
    res = model.forward_pairs(clouds)
    torch.cuda.synchronize(); print('forward_pairs', npairs, 'pairs: %.1f ms' % ((time.time() - t) * 1e3), 'mem GB', torch.cuda.max_memory_allocated() / 1e9)
print(res[0][0][:8], res[0][1][:8], res[0][2][:4])
# per-API timing
L = _lib.lib(); L.enabled = True
names = [k for k in _lib.KERNELS_PER_CALL]
L.reset(timed=names)
torch.cuda.synchronize(); t = time.time()
res = model.forward_pairs(clouds)
torch.cuda.synchronize(); tot = (time.time() - t) * 1e3
L.enabled = False
print('total %.1f ms, launches %d' % (tot, L.launches()))
rows = sorted(((L.timed_ms(n)[0], L.timed_ms(n)[1], n) for n in names), reverse=True)
for ms, cnt, n in rows:
    if cnt: print('%-32s %5d calls %9.2f ms' % (n, cnt, ms))
