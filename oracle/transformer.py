"""TEST INFRASTRUCTURE ONLY -- torch-CPU fp32 restatement of the reference's superpoint transformer and coarse
matching, functional and driven by a reference-format state_dict:

  sinusoidal_embedding          <- transformer/positional_embedding.py:8-34
  geometric_structure_embedding <- geotransformer/geotransformer.py:69-121 (pair distance + 3-NN triplet angles)
  rpe_self_attention_eq         <- transformer/rpe_transformer.py:39-131 (equivariant branch) + :134-194
  cross_attention_inv_eq        <- transformer/vanilla_transformer.py:39-85 (4-D value) + :872-946
  attention_output              <- transformer/output_layer.py:7-22
  geometric_transformer         <- geotransformer.py:213-317 + conditional_transformer.py:251-315
                                   (block lists made of 'self_eq' and 'cross', i.e. SE3ET-I / I2)
  superpoint_matching           <- geotransformer/superpoint_matching.py:13-55, ops/pairwise_distance.py:18-31
                                   canonical top-k order (score desc, flat index asc) per SURVEY 8(c)-iii

Parity pinned by tests/test_oracle_transformer.py against fixtures from the unmodified reference modules.
"""
import math

import torch
import torch.nn.functional as F

from .e2pn import Params


def sinusoidal_embedding(x, d_model):
    div = torch.exp(torch.arange(0, d_model, 2).float() * (-math.log(10000.0) / d_model))
    om = x[..., None] * div
    return torch.stack([torch.sin(om), torch.cos(om)], dim=-1).reshape(*x.shape, d_model)


def embedding_indices(points, sigma_d, sigma_a, k):
    """points (N,3) -> d_idx (N,N), a_idx (N,N,k)."""
    sq = (points ** 2).sum(-1)
    d2 = (sq[:, None] - 2.0 * points @ points.t() + sq[None, :]).clamp(min=0.0)
    dist = torch.sqrt(d2)
    knn = dist.topk(k=k + 1, dim=1, largest=False)[1][:, 1:]  # (N, k)
    ref_vec = points[knn] - points[:, None, :]  # (N, k, 3)
    anc_vec = points[None, :, :] - points[:, None, :]  # (N, N, 3): anc[n, m] = p_m - p_n
    ref_e = ref_vec[:, None, :, :].expand(-1, points.shape[0], -1, -1)
    anc_e = anc_vec[:, :, None, :].expand(-1, -1, k, -1)
    sin_v = torch.linalg.norm(torch.cross(ref_e, anc_e, dim=-1), dim=-1)
    cos_v = (ref_e * anc_e).sum(-1)
    ang = torch.atan2(sin_v, cos_v)
    return dist / sigma_d, ang * (180.0 / (sigma_a * math.pi))


def geometric_structure_embedding(p, points, d_model, sigma_d, sigma_a, k):
    d_idx, a_idx = embedding_indices(points, sigma_d, sigma_a, k)
    d_emb = F.linear(sinusoidal_embedding(d_idx, d_model), p["proj_d.weight"], p["proj_d.bias"])
    a_emb = F.linear(sinusoidal_embedding(a_idx, d_model), p["proj_a.weight"], p["proj_a.bias"]).amax(dim=2)
    return d_emb + a_emb  # (N, N, C)


def _lin(p, name, x):
    return F.linear(x, p[name + ".weight"], p[name + ".bias"])


def _ln(p, name, x):
    return F.layer_norm(x, (x.shape[-1],), p[name + ".weight"], p[name + ".bias"], 1e-5)


def attention_output(p, x):
    h = _lin(p, "squeeze", F.relu(_lin(p, "expand", x)))
    return _ln(p, "norm", x + h)


def rpe_self_attention_eq(p, x, emb, heads):
    """x: (A, N, C) equivariant states, emb: (N, N, C). One RPETransformerLayer (equivariant)."""
    a, n, c = x.shape
    hc = c // heads
    ap = p.sub("attention").sub("attention")
    q = _lin(ap, "proj_q", x).view(a, n, heads, hc).permute(0, 2, 1, 3)  # a h n c
    k = _lin(ap, "proj_k", x).view(a, n, heads, hc).permute(0, 2, 1, 3)
    v = _lin(ap, "proj_v", x).view(a, n, heads, hc).permute(0, 2, 1, 3)
    pe = _lin(ap, "proj_p", emb).view(n, n, heads, hc).permute(2, 0, 1, 3)  # h n m c
    s = (torch.einsum("ahnc,ahmc->ahnm", q, k) + torch.einsum("ahnc,hnmc->ahnm", q, pe)) / hc ** 0.5
    s = F.softmax(s, dim=-1)
    hid = torch.matmul(s, v).permute(0, 2, 1, 3).reshape(a, n, c)
    al = p.sub("attention")
    y = _ln(al, "norm", _lin(al, "linear", hid) + x)
    return attention_output(p.sub("output"), y)


def cross_attention_inv_eq(p, q_inv, k_inv, v_eq, heads):
    """q_inv (N, C), k_inv (M, C) invariant; v_eq (A, M, C) equivariant -> (A, N, C). One TransformerLayer."""
    n, c = q_inv.shape
    a, m, _ = v_eq.shape
    hc = c // heads
    ap = p.sub("attention").sub("attention")
    q = _lin(ap, "proj_q", q_inv).view(n, heads, hc).permute(1, 0, 2)
    k = _lin(ap, "proj_k", k_inv).view(m, heads, hc).permute(1, 0, 2)
    v = _lin(ap, "proj_v", v_eq).view(a, m, heads, hc).permute(0, 2, 1, 3)
    s = F.softmax(torch.einsum("hnc,hmc->hnm", q, k) / hc ** 0.5, dim=-1)
    hid = torch.matmul(s[None], v).permute(0, 2, 1, 3).reshape(a, n, c)
    al = p.sub("attention")
    y = _ln(al, "norm", _lin(al, "linear", hid) + q_inv[None])  # (N, C) residual broadcast over anchors
    return attention_output(p.sub("output"), y)


def geometric_transformer(sd, ref_points, src_points, ref_feats, src_feats, blocks, hidden_dim, heads, sigma_d,
                          sigma_a, angle_k, prefix="transformer."):
    """ref_feats (N, A, Cin), src_feats (M, A, Cin) -> (N, Cout), (M, Cout) [+ embeddings for inspection]."""
    p = Params(sd, prefix)
    emb0 = geometric_structure_embedding(p.sub("embedding"), ref_points, hidden_dim, sigma_d, sigma_a, angle_k)
    emb1 = geometric_structure_embedding(p.sub("embedding"), src_points, hidden_dim, sigma_d, sigma_a, angle_k)
    f0 = _lin(p, "in_proj", ref_feats.transpose(0, 1))  # (A, N, C)
    f1 = _lin(p, "in_proj", src_feats.transpose(0, 1))
    f0_eq = f1_eq = None
    for i, block in enumerate(blocks):
        lp = p.sub("transformer").sub("layers").sub(str(i))
        if block == "self_eq":
            f0 = rpe_self_attention_eq(lp, f0 if f0_eq is None else f0_eq, emb0, heads)
            f1 = rpe_self_attention_eq(lp, f1 if f1_eq is None else f1_eq, emb1, heads)
            f0_eq, f1_eq = f0, f1
            f0, f1 = f0.amax(0), f1.amax(0)
        elif block == "cross":
            # ref side first; the src side then attends to the UPDATED ref (conditional_transformer.py:297-302)
            f0_eq = cross_attention_inv_eq(lp, f0, f1, f1_eq, heads)
            f0 = f0_eq.amax(0)
            f1_eq = cross_attention_inv_eq(lp, f1, f0, f0_eq, heads)
            f1 = f1_eq.amax(0)
        else:
            raise NotImplementedError(block)
    return _lin(p, "out_proj", f0), _lin(p, "out_proj", f1), emb0, emb1


def matching_scores(ref_feats, src_feats, dual_normalization=True):
    s = torch.exp(-(2.0 - 2.0 * ref_feats @ src_feats.t()).clamp(min=0.0))
    if dual_normalization:
        s = (s / s.sum(dim=1, keepdim=True)) * (s / s.sum(dim=0, keepdim=True))
    return s


def canonical_topk(scores, k):
    """(score desc, flat index asc) -- the deterministic refinement of torch.topk's unspecified tie order."""
    flat = scores.reshape(-1)
    k = min(k, flat.numel())
    order = torch.argsort(-flat, stable=True)[:k]
    return flat[order], order


def superpoint_matching(ref_feats, src_feats, ref_masks, src_masks, num_correspondences, dual_normalization=True):
    ri = torch.nonzero(ref_masks, as_tuple=True)[0]
    si = torch.nonzero(src_masks, as_tuple=True)[0]
    s = matching_scores(ref_feats[ri], src_feats[si], dual_normalization)
    vals, idx = canonical_topk(s, num_correspondences)
    return ri[idx // s.shape[1]], si[idx % s.shape[1]], vals
