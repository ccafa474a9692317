import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
import helpers
from oracle import transformer as ot
from test_transformer_gpu import build_transformer
from test_oracle_transformer import coarse_inputs
g = np.load('/root/repo/tests/golden/model_small.npz')
S = helpers.SMALL_CFG
rp, sp, _, _ = coarse_inputs(g)
net, sd = build_transformer()
with torch.no_grad():
    e = net.embedding(rp[None].cuda())[0].cpu()
p = ot.Params(sd, "transformer.embedding.")
want = ot.geometric_structure_embedding(p, rp, S["hidden_dim"], S["sigma_d"], S["sigma_a"], S["angle_k"])
err = (e - want).abs()
print('max', err.max().item(), 'mean', err.mean().item(), 'want absmax', want.abs().max().item())
pair_err = err.amax(-1)
idx = torch.nonzero(pair_err > 0.03)
print('bad pairs', idx.shape[0], 'of', pair_err.numel(), idx[:10].tolist())
print('diag err', torch.diagonal(pair_err).max().item(), 'offdiag', (pair_err - torch.diag(torch.diagonal(pair_err))).max().item())
rows = torch.unique(idx[:,0]); print('bad rows', rows.tolist())
from se3et_b200.ops import transformer_ops as T
from se3et_b200.modules import transformer as MT
ctx = MT.CloudContext([len(rp)], [], 1, 1, 'cuda:0')
idx4 = T.geo_embed_indices(rp.cuda(), ctx.cu, ctx.max_n, ctx.eoff, ctx.R, 0.2, 15.0, 3).cpu().view(len(rp), len(rp), 4)
d, a = ot.embedding_indices(rp, 0.2, 15.0, 3)
print('gpu diag', idx4[range(6), range(6)].tolist())
print('cpu diag d', torch.diagonal(d)[:6].tolist())
print('cpu diag a', a[range(6), range(6)].tolist())
