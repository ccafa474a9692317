import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from se3et_b200 import synthetic, training as TR
from se3et_b200.model import create_model, make_cfg
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda')
cfg = make_cfg('se3eti.3dmatch')
torch.manual_seed(0)
model = create_model(cfg).to(dev).train()
opt = torch.optim.Adam(TR.trainable_parameters(model), lr=1e-4)
p = synthetic.make_3dmatch_pair(101)
rng = np.random.default_rng(0)
for _ in range(2):
    TR.training_step(model, p['ref_points'], p['src_points'], p['transform'], optimizer=opt, rng=rng)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    TR.training_step(model, p['ref_points'], p['src_points'], p['transform'], optimizer=opt, rng=rng)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
