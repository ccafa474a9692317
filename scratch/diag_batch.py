import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import helpers
from test_e2pn_gpu import _build_backbone, rel_err
from se3et_b200.precompute import precompute_data_stack_mode
DEV='cuda:0'
S = helpers.SMALL_CFG
net,_ = _build_backbone()
g = np.load('/root/repo/tests/golden/model_small.npz')
pts_a, lens_a = g['in_points'], g['in_lengths']
pts_b, lens_b = helpers.small_pair(index=12, crop=1.1)
def run(pts, lens, offs=False):
    d = precompute_data_stack_mode(torch.from_numpy(pts).to(DEV), torch.from_numpy(lens).to(DEV), 4, S['init_voxel'], S['init_radius'], [38,36,36,38])
    if offs: d['pair_offsets'] = [torch.stack([l.new_zeros(()), l[:2].sum(), l.sum()]) for l in d['lengths']]
    feats = {}
    hooks = []
    for name, m in net.named_children():
        hooks.append(m.register_forward_hook(lambda mod, inp, out, name=name: feats.__setitem__(name, out.float() if torch.is_tensor(out) else None)))
    with torch.no_grad():
        out = net(torch.ones(len(pts),1,device=DEV), d)
    for h in hooks: h.remove()
    return d, feats
d1, f1 = run(pts_a, lens_a)
d1b, f1b = run(pts_a, lens_a)
for k in f1:
    if f1[k] is not None: print('determinism', k, rel_err(f1[k], f1b[k]))
d2, f2 = run(np.concatenate([pts_a, pts_b]), np.concatenate([lens_a, lens_b]), True)
for k in f1:
    if f1[k] is None: continue
    n = f1[k].shape[0]
    print('batched vs single', k, n, rel_err(f2[k][:n], f1[k]))
print('--- pyramid compare')
for k in ('points','neighbors','subsampling','upsampling'):
    for i,(a,b) in enumerate(zip(d1[k], d2[k])):
        n=a.shape[0]; w=a.shape[1]
        bb=b[:n,:w]
        if k=='points': print(k,i,torch.equal(a,bb))
        else:
            pa = a>=d1['points'][i if k!='upsampling' else i+1].shape[0] if k!='subsampling' else a>=d1['points'][i].shape[0]
            ns2 = d2['points'][i+1].shape[0] if k=='upsampling' else d2['points'][i].shape[0]
            pb = bb>=ns2
            print(k,i,a.shape,b.shape, torch.equal(pa,pb), torch.equal(torch.where(pa,0,a), torch.where(pb,0,bb)), bool((b[:n,w:]>=ns2).all()))
# inside encoder2_1
blk = net.encoder2_1
def run_blk(d, x, offs):
    segs = d.get('pair_offsets',[None]*4)
    res={}
    hooks=[m.register_forward_hook(lambda mod,inp,out,name=name: res.__setitem__(name, out.float() if torch.is_tensor(out) else None)) for name,m in blk.named_modules() if name in ('unary1','interso3.conv')]
    with torch.no_grad():
        out = blk(x, d['points'][1], d['points'][0], d['subsampling'][0], seg=segs[1], s_seg=segs[0])
    for h in hooks: h.remove()
    res['out']=out.float()
    return res
r1 = run_blk(d1, f1['encoder1_2'].bfloat16(), False)
r2 = run_blk(d2, f2['encoder1_2'].bfloat16(), True)
for k in r1:
    n=r1[k].shape[0]
    print('enc2_1', k, rel_err(r2[k][:n], r1[k]))
