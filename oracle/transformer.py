"""TEST INFRASTRUCTURE ONLY -- torch-CPU fp32 restatement of the reference's superpoint transformer and coarse
matching, functional and driven by a reference-format state_dict:

  sinusoidal_embedding          <- transformer/positional_embedding.py:8-34
  geometric_structure_embedding <- geotransformer/geotransformer.py:69-121 (pair distance + 3-NN triplet angles)
  rpe_self_attention_eq         <- transformer/rpe_transformer.py:39-131 (equivariant branch) + :134-194
  cross_attention_inv_eq        <- transformer/vanilla_transformer.py:39-85 (4-D value) + :872-946
  attention_output              <- transformer/output_layer.py:7-22
  geometric_transformer         <- geotransformer.py:213-317 + conditional_transformer.py:251-390
                                   (block lists of SE3ET-I / I2: 'self_eq', 'cross'; and of SE3ET-E / E2:
                                   + 'cross_a_soft', 'cross_r_soft', 'self', invariant 'cross')
  cross_attention_eq            <- transformer/vanilla_transformer.py:247-476, 509-575, 751-870 (a_soft / r_soft)
  rot_compress, eq2inv_soft     <- transformer/output_layer.py:24-47, conditional_transformer.py:209-249
  sh_equiv_embedding            <- geotransformer.py:40-67 with the e3nn convention stated there (e3nn is NOT in
                                   this image and the reference does not pin its version: PARITY UNPINNED for
                                   this one term; everything else is pinned through fixtures)
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


def octahedral_rotation_perms():
    """(24, 6) int64: row r = the permutation of the octahedron vertices [+z,+x,+y,-x,-y,-z] under the r-th proper
    rotation, perm[r][a] = index of R_r v_a (what fr.get_relativeV_index calls trace_idx_ori, rotation.py:581-601).
    The ORDER of the rows is the reference's only up to relabelling; every quantity of the forward pass sums over r."""
    import itertools
    import numpy as np
    vs = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=np.float64)
    rows = []
    for perm in itertools.permutations(range(3)):
        for signs in itertools.product((1, -1), repeat=3):
            R = np.zeros((3, 3))
            for i in range(3):
                R[i, perm[i]] = signs[i]
            if np.linalg.det(R) > 0:
                rows.append([int(np.argmin(((vs - R @ v) ** 2).sum(1))) for v in vs])
    rows.sort()
    return torch.tensor(rows, dtype=torch.int64)


def sh_equiv_embedding(points, anchors):
    """geotransformer.py:57-67 with n_level_equiv = 2: real spherical harmonics l = 0, 1 of p_n - p_m
    (`normalize=True`, e3nn 'integral' normalisation, l = 1 identified with (x, y, z)) rotated by the Wigner-D of
    anchors^T, which for l = 1 is the matrix itself.  points (N, 3), anchors (A, 3, 3) -> (A, N, N, 4).
    ASSUMED CONVENTION (e3nn absent, unpinned): Y0 = 1 / (2 sqrt(pi)), Y1 = sqrt(3 / (4 pi)) * unit(p_n - p_m)."""
    diff = points[:, None, :] - points[None, :, :]
    unit = F.normalize(diff, dim=-1)
    y0 = torch.full(diff.shape[:2] + (1,), 0.5 / math.sqrt(math.pi), dtype=points.dtype)
    y1 = math.sqrt(3.0 / (4.0 * math.pi)) * unit
    d1 = anchors.transpose(1, 2)  # D^1(anchors^T) = anchors^T
    y1a = torch.einsum("acd,nmd->anmc", d1, y1)
    return torch.cat([y0[None].expand(anchors.shape[0], -1, -1, -1), y1a], dim=-1)


def rpe_self_attention(p, x, emb, heads):
    """Non-equivariant RPETransformerLayer ('self' blocks): x (N, C), emb (N, N, C)."""
    return rpe_self_attention_eq(p, x[None], emb, heads)[0]


def rpe_self_attention_eq_sh(p, x, emb, emb_eq, heads):
    """rpe_self_attention_eq plus the q . proj_eq(embed_eq) score term (rpe_transformer.py:76-79, 94-95)."""
    a, n, c = x.shape
    hc = c // heads
    ap = p.sub("attention").sub("attention")
    q = _lin(ap, "proj_q", x).view(a, n, heads, hc).permute(0, 2, 1, 3)
    k = _lin(ap, "proj_k", x).view(a, n, heads, hc).permute(0, 2, 1, 3)
    v = _lin(ap, "proj_v", x).view(a, n, heads, hc).permute(0, 2, 1, 3)
    pe = _lin(ap, "proj_p", emb).view(n, n, heads, hc).permute(2, 0, 1, 3)
    eq = _lin(ap, "proj_eq", emb_eq).view(a, n, n, heads, hc).permute(0, 3, 1, 2, 4)  # a h n m c
    s = (torch.einsum("ahnc,ahmc->ahnm", q, k) + torch.einsum("ahnc,hnmc->ahnm", q, pe) +
         torch.einsum("ahnc,ahnmc->ahnm", q, eq)) / hc ** 0.5
    s = F.softmax(s, dim=-1)
    hid = torch.matmul(s, v).permute(0, 2, 1, 3).reshape(a, n, c)
    al = p.sub("attention")
    y = _ln(al, "norm", _lin(al, "linear", hid) + x)
    return attention_output(p.sub("output"), y)


def anchor_mixing_weights(g, mode, perms):
    """g (A, A) = pooled non-negative anchor-pair scores; -> (W (A, A), attn_r (R,) or None) with
    hidden[a] = sum_e W[a, e] softmax(S[a, e]) v[e]   (vanilla_transformer.py:466-476 a_soft; 509-575, 839-845 r_soft:
    brahnm scores weighted by attn_r and gathered through trace_idx_ori collapse to the same form)."""
    if mode == "a_soft":
        return g / g.sum(1, keepdim=True), None
    a = g.shape[0]
    attn_ar = g[torch.arange(a)[None, :], perms]  # (R, A): g[a, perms[r][a]]
    attn_r = attn_ar.mean(1)
    attn_r = attn_r / attn_r.sum()
    w = torch.zeros_like(g)
    for r in range(perms.shape[0]):
        w[torch.arange(a), perms[r]] += attn_r[r]
    return w, attn_r


def cross_attention_eq(p, x_q, x_k, heads, mode, perms, positive="sq"):
    """One equivariant TransformerLayer with MultiHeadAttentionEQ in mode a_soft / r_soft.
    x_q (A, N, C), x_k (A, M, C) -> (out (A, N, C), W (A, A), attn_r)."""
    a, n, c = x_q.shape
    m = x_k.shape[1]
    hc = c // heads
    ap = p.sub("attention").sub("attention")
    q = _lin(ap, "proj_q", x_q).view(a, n, heads, hc).permute(0, 2, 1, 3)
    k = _lin(ap, "proj_k", x_k).view(a, m, heads, hc).permute(0, 2, 1, 3)
    v = _lin(ap, "proj_v", x_k).view(a, m, heads, hc).permute(0, 2, 1, 3)
    s = torch.einsum("ahnc,ehmc->aehnm", q, k) / hc ** 0.5
    g = s.mean(2)
    if positive == "sq":
        g = g ** 2
    elif positive == "softplus":
        g = F.softplus(g)
    elif positive == "sigmoid":
        g = torch.sigmoid(g)
    elif positive == "relu":
        g = F.relu(g)
    elif positive == "abs":
        g = g.abs()
    else:
        raise NotImplementedError(positive)
    g = g.mean((-2, -1))  # (a, e)
    w, attn_r = anchor_mixing_weights(g, mode, perms)
    prob = F.softmax(s, dim=-1) * w[:, :, None, None, None]
    hid = torch.einsum("aehnm,ehmc->ahnc", prob, v).permute(0, 2, 1, 3).reshape(a, n, c)
    al = p.sub("attention")
    y = _ln(al, "norm", _lin(al, "linear", hid) + x_q)
    return attention_output(p.sub("output"), y), w, attn_r


def cross_attention_inv(p, q_inv, k_inv, heads):
    """Invariant TransformerLayer ('cross' after the features were pooled): (N, C) x (M, C) -> (N, C)."""
    return cross_attention_inv_eq(p, q_inv, k_inv, k_inv[None], heads)[0]


def rot_compress(p, x):
    """RotCompressOutput (output_layer.py:24-47): x (A, N, C) -> (N, C)."""
    a, n, c = x.shape
    h = _lin(p, "squeeze", F.relu(_lin(p, "expand", x.permute(1, 0, 2).reshape(n, a * c))))
    return _ln(p, "norm", x.amax(0) + h)


def geometric_transformer_eq(sd, ref_points, src_points, ref_feats, src_feats, blocks, hidden_dim, heads, sigma_d,
                             sigma_a, angle_k, anchors, n_level_equiv=2, positive="sq", prefix="transformer."):
    """SE3ET-E / E2 block lists (experiments/se3ete.3dmatch/config.py:194): equivariant states until the
    'cross_r_soft' block, soft alignment + RotCompressOutput (align_mode '0'), invariant blocks afterwards."""
    p = Params(sd, prefix)
    perms = octahedral_rotation_perms()
    emb0 = geometric_structure_embedding(p.sub("embedding"), ref_points, hidden_dim, sigma_d, sigma_a, angle_k)
    emb1 = geometric_structure_embedding(p.sub("embedding"), src_points, hidden_dim, sigma_d, sigma_a, angle_k)
    eq0 = eq1 = None
    if n_level_equiv > 0:
        assert n_level_equiv == 2
        eq0, eq1 = sh_equiv_embedding(ref_points, anchors), sh_equiv_embedding(src_points, anchors)
    f0 = _lin(p, "in_proj", ref_feats.transpose(0, 1))  # (A, N, C)
    f1 = _lin(p, "in_proj", src_feats.transpose(0, 1))
    tp = p.sub("transformer")
    for i, block in enumerate(blocks):
        lp = tp.sub("layers").sub(str(i))
        if block == "self_eq":
            if eq0 is not None:
                f0 = rpe_self_attention_eq_sh(lp, f0, emb0, eq0, heads)
                f1 = rpe_self_attention_eq_sh(lp, f1, emb1, eq1, heads)
            else:
                f0 = rpe_self_attention_eq(lp, f0, emb0, heads)
                f1 = rpe_self_attention_eq(lp, f1, emb1, heads)
        elif block in ("cross_a_soft", "cross_r_soft"):
            mode = block[len("cross_"):]
            f0, w0, _ = cross_attention_eq(lp, f0, f1, heads, mode, perms, positive)
            f1, w1, _ = cross_attention_eq(lp, f1, f0, heads, mode, perms, positive)
            if mode == "r_soft" and i + 1 < len(blocks) and "_" not in blocks[i + 1]:
                # eq2inv_soft, align_mode '0': the source anchors are mixed by the reference side's rotation weights
                f1 = torch.einsum("ae,enc->anc", w0, f1)
                f0, f1 = rot_compress(tp.sub("rotcompress"), f0), rot_compress(tp.sub("rotcompress"), f1)
        elif block == "self":
            f0, f1 = rpe_self_attention(lp, f0, emb0, heads), rpe_self_attention(lp, f1, emb1, heads)
        elif block == "cross":
            f0 = cross_attention_inv(lp, f0, f1, heads)
            f1 = cross_attention_inv(lp, f1, f0, heads)
        else:
            raise NotImplementedError(block)
    return _lin(p, "out_proj", f0), _lin(p, "out_proj", f1)


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
