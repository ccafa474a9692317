"""TEST INFRASTRUCTURE ONLY -- torch-CPU fp32 restatement of the reference's E2PN backbone math.

Functional (no nn.Module), driven by a reference-format state_dict, written from the algorithm:

  octahedral_tables        <- blocks_epn.py:111-332 (init_KP / init_anchors / init_permute_idxs_*),
                              utils_epn/rotation.py:484-523, anchors.py:41-44,85-90
  kpconv_inter_so3         <- blocks_epn.py:454-546 + 334-390 (rot_by_permute, non_sep_conv, 'linear', 'sum')
  group_norm_epn           <- blocks_epn.py:684-701 (GroupNorm over (C/G) x A x N, eps 1e-5)
  unary_block_epn          <- blocks_epn.py:639-665
  simple_block_epn         <- blocks_epn.py:770-796 (+ KPConvInterSO3Block :703-743)
  resnet_bottleneck_epn    <- blocks_epn.py:798-852, max_pool e2pn/blocks.py:93-110
  e2pn_forward             <- experiments/se3eti.3dmatch/backbone.py:35-77 (4 stages) and
                              experiments/se3eti.kitti/backbone.py:41-99 (5 stages)
  group_norm / unary_block <- kpconv/modules.py:33-101, nearest_upsample kpconv/functional.py:6-22

Parity pinned by tests/test_oracle_e2pn.py against fixtures produced by importing the unmodified reference
modules (tests/golden/make_model_golden.py).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# anchors of get_anchorsV24(): the 6 representative rotations of the octahedral group (one per vertex);
# SURVEY.md 8(c) golden constants, exact entries in {-1, 0, 1}
_ANCHORS = np.array([
    [[1, 0, 0], [0, 1, 0], [0, 0, 1]],
    [[0, 0, 1], [0, 1, 0], [-1, 0, 0]],
    [[0, -1, 0], [0, 0, 1], [-1, 0, 0]],
    [[0, 0, -1], [0, -1, 0], [-1, 0, 0]],
    [[0, 1, 0], [0, 0, -1], [-1, 0, 0]],
    [[1, 0, 0], [0, -1, 0], [0, 0, -1]],
], dtype=np.float64)

_VERTS = np.array([[0, 0, 1], [1, 0, 0], [0, 1, 0], [-1, 0, 0], [0, -1, 0], [0, 0, -1]], dtype=np.float64)
_FACES = np.array([[1, 1, 1], [-1, 1, 1], [-1, -1, 1], [1, -1, 1], [1, 1, -1], [-1, 1, -1], [-1, -1, -1], [1, -1, -1]],
                  dtype=np.float64) / math.sqrt(3.0)


def unit_kernel_points():
    """15 kernel directions: 6 octahedron vertices, 8 face normals, centre (blocks_epn.py:161-170)."""
    return np.concatenate([_VERTS, _FACES, np.zeros((1, 3))], 0)


def octahedral_tables():
    """Returns dict(anchors (6,3,3), quotient (4,3,3), kp_unit (15,3), kidx (15,6), ridx (6,6), k_real)."""
    anchors = _ANCHORS
    ang = np.arange(4) * (np.pi / 2)
    quotient = np.stack([np.array([[np.cos(t), -np.sin(t), 0], [np.sin(t), np.cos(t), 0], [0, 0, 1]]) for t in ang])
    kp = unit_kernel_points()
    K = kp.shape[0]
    # weight-sharing classes: orbits of the kernel points under the quotient (z-rotation) subgroup
    cls = -np.ones(K, dtype=np.int64)
    n_cls = 0
    for i in range(K):
        if cls[i] >= 0:
            continue
        for q in quotient:
            j = int(np.argmin(np.linalg.norm(kp - q @ kp[i], axis=1)))
            cls[j] = n_cls
        n_cls += 1
    # kidx[k, r]: class of the kernel point that anchor r rotates ONTO k
    kidx = np.zeros((K, 6), dtype=np.int64)
    for r in range(6):
        rot = kp @ anchors[r].T  # rot[k2] = R_r kp[k2]
        for k in range(K):
            k2 = int(np.argmin(np.linalg.norm(rot - kp[k], axis=1)))
            kidx[k, r] = cls[k2]
    # ridx[a, r]: anchor b with R_r R_b in the coset of R_a (R_a * quotient)
    ridx = np.zeros((6, 6), dtype=np.int64)
    for a in range(6):
        for r in range(6):
            best, best_cos = 0, -2.0
            for b in range(6):
                prod = anchors[r] @ anchors[b]
                c = max(0.5 * (np.trace((anchors[a] @ q).T @ prod) - 1.0) for q in quotient)
                if c > best_cos + 1e-9:
                    best, best_cos = b, c
            ridx[a, r] = best
    return {"anchors": anchors, "quotient": quotient, "kp_unit": kp, "kidx": kidx, "ridx": ridx, "k_real": n_cls}


def leaky(x):
    return F.leaky_relu(x, 0.1)


def group_norm_epn(x, num_groups, weight, bias):
    """x: (N, A, C). Statistics over (C/G) x A x N, i.e. the whole stacked ref+src tensor."""
    n, a, c = x.shape
    g = x.reshape(n * a, num_groups, c // num_groups)
    mean = g.mean(dim=(0, 2), keepdim=True)
    var = g.var(dim=(0, 2), unbiased=False, keepdim=True)
    y = ((g - mean) / torch.sqrt(var + 1e-5)).reshape(n, a, c)
    return y * weight + bias


def group_norm(x, num_groups, weight, bias):
    """x: (N, C) invariant features (kpconv/modules.py:33-50)."""
    return group_norm_epn(x[:, None, :], num_groups, weight, bias)[:, 0, :]


def kpconv_inter_so3(q_pts, s_pts, neighb_inds, x, weights, kernel_points, kp_extent, kidx, ridx):
    """x: (Ns, A, Cin); weights: (K_real, A, Cin, Cout) -> (Nq, A, Cout)."""
    ns = s_pts.shape[0]
    s_pad = torch.cat([s_pts, torch.full((1, 3), 1e6, dtype=s_pts.dtype)], 0)
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)
    rel = s_pad[neighb_inds] - q_pts[:, None, :]  # (P, H, 3)
    dist = torch.sqrt(((rel[:, :, None, :] - kernel_points[None, None]) ** 2).sum(-1))  # (P, H, K)
    infl = torch.clamp(1.0 - dist / kp_extent, min=0.0)
    nx = x_pad[neighb_inds]  # (P, H, A, Cin)
    wf = torch.einsum("pnac,pnk->kpac", nx, infl)  # (K, P, A, Cin)
    kidx_t = torch.as_tensor(kidx, dtype=torch.long)
    ridx_t = torch.as_tensor(ridx, dtype=torch.long)
    # W_eff[k, a, r] = W[kidx[k, r], ridx[a, r]]
    w_eff = weights[kidx_t[:, None, :], ridx_t[None, :, :]]  # (K, A, R, Cin, Cout)
    del ns
    return torch.einsum("kpac,karcd->prd", wf, w_eff)


def max_pool(x, inds):
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)
    return x_pad[inds].amax(dim=1)


def nearest_upsample(x, inds):
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)
    return x_pad[inds[:, 0]]


class Params:
    """Reads tensors of one sub-module out of a flat reference state_dict."""

    def __init__(self, sd, prefix=""):
        self.sd, self.prefix = sd, prefix

    def sub(self, name):
        return Params(self.sd, self.prefix + name + ".")

    def __getitem__(self, name):
        return self.sd[self.prefix + name].float()

    def has(self, name):
        return (self.prefix + name) in self.sd


def unary_block_epn(p, x, groups, relu=True):
    y = F.linear(x, p["mlp.weight"], p["mlp.bias"])
    y = group_norm_epn(y, groups, p["norm.norm.weight"], p["norm.norm.bias"])
    return leaky(y) if relu else y


def _conv(p, x, q_pts, s_pts, inds, tabs):
    return kpconv_inter_so3(q_pts, s_pts, inds, x, p["weights"], p["kernel_points"], tabs["extent"], tabs["kidx"],
                            tabs["ridx"])


def interso3_block(p, x, q_pts, s_pts, inds, groups, tabs):
    y = _conv(p.sub("conv"), x, q_pts, s_pts, inds, tabs)
    return leaky(group_norm_epn(y, groups, p["norm.norm.weight"], p["norm.norm.bias"]))


def simple_block_epn(p, x, q_pts, s_pts, inds, groups, tabs):
    y = interso3_block(p.sub("interso3"), x, q_pts, s_pts, inds, groups, tabs)
    return leaky(group_norm_epn(y, groups, p["norm.norm.weight"], p["norm.norm.bias"]))


def resnet_bottleneck_epn(p, x, q_pts, s_pts, inds, groups, tabs, strided):
    skip = x
    y = unary_block_epn(p.sub("unary1"), x, groups) if p.has("unary1.mlp.weight") else x
    y = interso3_block(p.sub("interso3"), y, q_pts, s_pts, inds, groups, tabs)
    y = leaky(group_norm_epn(y, groups, p["norm.norm.weight"], p["norm.norm.bias"]))
    y = unary_block_epn(p.sub("unary2"), y, groups, relu=False)
    if strided:
        skip = max_pool(skip, inds)
    if p.has("skip_conv.mlp.weight"):
        skip = unary_block_epn(p.sub("skip_conv"), skip, groups, relu=False)
    return leaky(y + skip)


def e2pn_forward(sd, feats, data_dict, init_sigma, groups, prefix="backbone.", tables=None):
    """4- or 5-stage E2PN. Returns [feats_f, latent_mid, feats_c] like the reference (feats_list reversed)."""
    t = tables or octahedral_tables()
    p = Params(sd, prefix)
    pts, nb, sub, up = data_dict["points"], data_dict["neighbors"], data_dict["subsampling"], data_dict["upsampling"]
    pts = [torch.as_tensor(v).float() for v in pts]
    nb = [torch.as_tensor(v).long() for v in nb]
    sub = [torch.as_tensor(v).long() for v in sub]
    up = [torch.as_tensor(v).long() for v in up]
    stages = len(pts)

    def tabs(level):
        return {"kidx": t["kidx"], "ridx": t["ridx"], "extent": init_sigma * (2 ** level)}

    x = feats.float()[:, None, :].expand(-1, 6, -1)  # LiftBlockEPN
    x = simple_block_epn(p.sub("encoder1_1"), x, pts[0], pts[0], nb[0], groups, tabs(0))
    x = resnet_bottleneck_epn(p.sub("encoder1_2"), x, pts[0], pts[0], nb[0], groups, tabs(0), False)
    inv = {}
    feats_s = {1: x}
    for s in range(2, stages + 1):
        lvl = s - 1
        x = resnet_bottleneck_epn(p.sub("encoder%d_1" % s), x, pts[lvl], pts[lvl - 1], sub[lvl - 1], groups,
                                  tabs(lvl - 1), True)
        x = resnet_bottleneck_epn(p.sub("encoder%d_2" % s), x, pts[lvl], pts[lvl], nb[lvl], groups, tabs(lvl), False)
        x = resnet_bottleneck_epn(p.sub("encoder%d_3" % s), x, pts[lvl], pts[lvl], nb[lvl], groups, tabs(lvl), False)
        feats_s[s] = x
        inv[s] = x.amax(dim=1)  # InvOutBlockEPN
    out = [x]
    latent = inv[stages]
    for s in range(stages - 1, 1, -1):
        latent = torch.cat([nearest_upsample(latent, up[s - 1]), inv[s]], dim=1)
        dp = p.sub("decoder%d" % s)
        latent = F.linear(latent, dp["mlp.weight"], dp["mlp.bias"])
        if dp.has("norm.norm.weight"):  # UnaryBlock; the last decoder is a LastUnaryBlock (no norm / activation)
            latent = leaky(group_norm(latent, groups, dp["norm.norm.weight"], dp["norm.norm.bias"]))
        out.append(latent)
    out.reverse()
    return out
