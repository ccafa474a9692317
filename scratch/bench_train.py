import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from se3et_b200 import synthetic, training as TR
from se3et_b200.model import create_model, make_cfg
dev = torch.device('cuda')
cfg = make_cfg('se3eti.3dmatch')
torch.manual_seed(0)
model = create_model(cfg).to(dev).train()
opt = torch.optim.Adam(TR.trainable_parameters(model), lr=1e-4)
p = synthetic.make_3dmatch_pair(101)
for mode in (None, torch.bfloat16):
    TR.RECOMPUTE['autocast'] = mode
    rng = np.random.default_rng(0)
    for _ in range(2):
        out = TR.training_step(model, p['ref_points'], p['src_points'], p['transform'], optimizer=opt, rng=rng)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        out = TR.training_step(model, p['ref_points'], p['src_points'], p['transform'], optimizer=opt, rng=rng)
    torch.cuda.synchronize()
    print('recompute', mode, '%.1f ms per step' % ((time.perf_counter() - t0) / 3 * 1e3), out, 'peak mem GB %.1f' % (torch.cuda.max_memory_allocated() / 1e9), flush=True)
