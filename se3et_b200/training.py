"""Training step of SE3ET-I (BASELINE.json configs[4]): forward on the CUDA path, backward by recomputation through
ATen, the reference's losses, one NCCL all-reduce of the flattened gradients.

    forward   GeoTransformer's own CUDA kernels (se3et_b200/model.py), under no_grad, inside torch.autograd.Function
    backward  the same computation restated with differentiable ATen ops over the SAME nn.Module parameters
              (`aten_backbone`, `aten_transformer`, `aten_log_optimal_transport` below), re-run under enable_grad and
              differentiated by autograd: "recompute-through-ATen" at whole-stage granularity.  Hand-written backward
              kernels are the next step; this first cut makes the training configuration runnable and
              measurable at 1/2/4/8 GPUs with the gradient exchange the reference's trainer performs.
    losses    CoarseMatchingLoss (weighted circle loss) + FineMatchingLoss (experiments/se3eti.3dmatch/loss.py:15-77,
              modules/loss/circle_loss.py:44-87), ground-truth superpoint correspondences from
              get_node_correspondences (modules/registration/matching.py:231-315), target sampling from
              SuperPointTargetGenerator (modules/geotransformer/superpoint_target.py:6-46)
    exchange  what DistributedDataParallel does in the reference (engine/base_trainer.py:181-189): gradients averaged
              over ranks, here as ONE flattened bucket per step (`allreduce_gradients`)

The reference trains one pair per GPU per iteration (batch_size 1, epoch_based_trainer.py:82-141); so does this.
Everything here is SE3ET-I ('self_eq' / 'cross' blocks); other variants raise."""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .modules import e2pn as ME
from .modules.sinkhorn import log_optimal_transport
from .ops.partition_ops import point_to_node_partition_stacked
from .precompute import precompute_data_stack_mode


# ======================================================================================================================
# ATen restatement of the path (differentiable; fp32)
# ======================================================================================================================
def _leaky(x):
    return F.leaky_relu(x, 0.1)


# dtype of the matrix products of the backward recomputation: None = fp32 (what the gradient-parity test checks);
# torch.bfloat16 = the precision the CUDA forward runs in (bf16 operands, fp32 accumulation; statistics stay fp32)
RECOMPUTE = {'autocast': None}


def _group_norm(x, gn):
    """GroupNormEPN / GroupNorm (blocks_epn.py:684-701): statistics over all rows (points x anchors) of a group."""
    x = x.float()
    shape, c, g = x.shape, x.shape[-1], gn.num_groups
    y = x.reshape(-1, g, c // g)
    mean = y.mean(dim=(0, 2), keepdim=True)
    var = y.var(dim=(0, 2), unbiased=False, keepdim=True)
    y = ((y - mean) * torch.rsqrt(var + gn.eps)).reshape(shape)
    return y * gn.weight + gn.bias


class _GatherRows(torch.autograd.Function):
    """x[idx] with an atomic scatter-add backward (index_add_): autograd's own backward of advanced indexing sorts the
    indices and serialises duplicates -- with ~38 references per support row it took 90 % of the first version's step."""

    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.rows = x.shape[0]
        return x[idx]

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        flat = g.reshape((idx.numel(),) + g.shape[idx.dim():])
        out = flat.new_zeros((ctx.rows,) + flat.shape[1:])
        out.index_add_(0, idx.reshape(-1), flat)
        return out, None


def gather_rows(x, idx):
    return _GatherRows.apply(x, idx)


def _kpconv(conv, q_pts, s_pts, idx, x):
    """KPConvInterSO3.forward (blocks_epn.py:454-546): x (Ns, A, Cin) -> (Nq, A, Cout)."""
    s_pad = torch.cat([s_pts, torch.full((1, 3), 1e6, dtype=s_pts.dtype, device=s_pts.device)], 0)
    x_pad = torch.cat([x, torch.zeros_like(x[:1])], 0)
    rel = s_pad[idx] - q_pts[:, None, :]                                                     # (P, H, 3)
    dist = torch.sqrt(((rel[:, :, None, :] - conv.kernel_points[None, None]) ** 2).sum(-1))   # (P, H, K)
    infl = torch.clamp(1.0 - dist / conv.KP_extent, min=0.0)
    wf = torch.einsum("pnac,pnk->kpac", gather_rows(x_pad, idx), infl)                        # (K, P, A, Cin)
    kidx, ridx = conv.kidx_rot[:, 0, :], conv.ridx_rot[0]                                     # (K, R), (A, R)
    w_eff = conv.weights[kidx[:, None, :], ridx[None, :, :]]                                  # (K, A, R, Cin, Cout)
    return torch.einsum("kpac,karcd->prd", wf, w_eff)


def _unary_epn(m, x, relu=True):
    y = _group_norm(F.linear(x, m.mlp.weight, m.mlp.bias), m.norm.norm)
    return _leaky(y) if relu else y


def _interso3(m, x, q_pts, s_pts, idx):
    return _leaky(_group_norm(_kpconv(m.conv, q_pts, s_pts, idx, x), m.norm.norm))


def _simple(m, x, q_pts, s_pts, idx):
    return _leaky(_group_norm(_interso3(m.interso3, x, q_pts, s_pts, idx), m.norm.norm))


def _resnet(m, x, q_pts, s_pts, idx):
    skip = x
    y = _unary_epn(m.unary1, x) if isinstance(m.unary1, ME.UnaryBlockEPN) else x
    y = _leaky(_group_norm(_interso3(m.interso3, y, q_pts, s_pts, idx), m.norm.norm))
    y = _unary_epn(m.unary2, y, relu=False)
    if 'strided' in m.block_name:
        skip = gather_rows(torch.cat([skip, torch.zeros_like(skip[:1])], 0), idx).amax(dim=1)
    if isinstance(m.skip_conv, ME.UnaryBlockEPN):
        skip = _unary_epn(m.skip_conv, skip, relu=False)
    return _leaky(y + skip)


def aten_backbone(bb, feats, dd):
    """E2PN.forward (experiments/se3eti.3dmatch/backbone.py:35-77) -> [feats_f, ..., feats_c (N_c, A, C)]."""
    pts, nb, sub, up = dd['points'], dd['neighbors'], dd['subsampling'], dd['upsampling']
    x = feats.float()[:, None, :].expand(-1, 6, -1)
    x = _simple(bb.encoder1_1, x, pts[0], pts[0], nb[0])
    x = _resnet(bb.encoder1_2, x, pts[0], pts[0], nb[0])
    inv = {}
    for s in range(2, bb.num_stages + 1):
        l = s - 1
        x = _resnet(getattr(bb, 'encoder%d_1' % s), x, pts[l], pts[l - 1], sub[l - 1])
        x = _resnet(getattr(bb, 'encoder%d_2' % s), x, pts[l], pts[l], nb[l])
        x = _resnet(getattr(bb, 'encoder%d_3' % s), x, pts[l], pts[l], nb[l])
        inv[s] = x.amax(dim=1)
    out = [x]
    latent = inv[bb.num_stages]
    for s in range(bb.num_stages - 1, 1, -1):
        lat_pad = torch.cat([latent, torch.zeros_like(latent[:1])], 0)
        latent = torch.cat([gather_rows(lat_pad, up[s - 1][:, 0].contiguous()), inv[s]], dim=1)
        dec = getattr(bb, 'decoder%d' % s)
        latent = F.linear(latent, dec.mlp.weight, dec.mlp.bias)
        if hasattr(dec, 'norm'):
            latent = _leaky(_group_norm(latent, dec.norm.norm))
        out.append(latent)
    out.reverse()
    return out


def _geo_embed(em, pts):
    """GeometricStructureEmbedding.forward (geotransformer.py:69-121) for one cloud: (n, 3) -> (n, n, C)."""
    with torch.no_grad():
        d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
        dist = torch.sqrt(d2.clamp_min(0))
        d_idx = dist / em.sigma_d
        k = em.angle_k
        knn = dist.topk(k=k + 1, dim=1, largest=False)[1][:, 1:]            # (n, k)
        ref_v = pts[knn] - pts[:, None, :]                                  # (n, k, 3)
        anc_v = pts[None, :, :] - pts[:, None, :]                           # (n, n, 3)
        ref_v = ref_v[:, None, :, :].expand(-1, pts.shape[0], -1, -1)
        anc_v = anc_v[:, :, None, :].expand(-1, -1, k, -1)
        sin_v = torch.linalg.norm(torch.cross(ref_v, anc_v, dim=-1), dim=-1)
        cos_v = (ref_v * anc_v).sum(-1)
        a_idx = torch.atan2(sin_v, cos_v) * em.factor_a                     # (n, n, k)
        div = em.embedding.div_term

        def sinus(u):
            om = u[..., None] * div
            return torch.stack([torch.sin(om), torch.cos(om)], dim=-1).flatten(-2)
        e_d, e_a = sinus(d_idx), sinus(a_idx)
    d_emb = F.linear(e_d, em.proj_d.weight, em.proj_d.bias)
    a_emb = F.linear(e_a, em.proj_a.weight, em.proj_a.bias).amax(dim=2)
    return d_emb + a_emb


def _attention_output(o, x):
    h = F.linear(F.relu(F.linear(x, o.expand.weight, o.expand.bias)), o.squeeze.weight, o.squeeze.bias)
    return F.layer_norm(x + h, (x.shape[-1],), o.norm.weight, o.norm.bias, o.norm.eps)


def _rpe_layer(layer, x, emb):
    """RPETransformerLayer, equivariant branch (rpe_transformer.py:56-194): x (n, A, C), emb (n, n, C)."""
    al = layer.attention
    at = al.attention
    n, a, c = x.shape
    h, hc = at.num_heads, at.d_model_per_head
    q = F.linear(x, at.proj_q.weight, at.proj_q.bias).view(n, a, h, hc)
    k = F.linear(x, at.proj_k.weight, at.proj_k.bias).view(n, a, h, hc)
    v = F.linear(x, at.proj_v.weight, at.proj_v.bias).view(n, a, h, hc)
    p = F.linear(emb, at.proj_p.weight, at.proj_p.bias).view(n, n, h, hc)
    s = (torch.einsum('nahc,mahc->ahnm', q, k) + torch.einsum('nahc,nmhc->ahnm', q, p)) / hc ** 0.5
    hid = torch.einsum('ahnm,mahc->nahc', F.softmax(s, dim=-1), v).reshape(n, a, c)
    y = F.linear(hid, al.linear.weight, al.linear.bias)
    y = F.layer_norm(y + x, (c,), al.norm.weight, al.norm.bias, al.norm.eps)
    return _attention_output(layer.output, y)


def _cross_layer(layer, q_inv, k_inv, v_eq):
    """TransformerLayer with invariant q / k and an equivariant value (vanilla_transformer.py:58-85, 872-946)."""
    al = layer.attention
    at = al.attention
    n, c = q_inv.shape
    m, a = v_eq.shape[0], v_eq.shape[1]
    h, hc = at.num_heads, at.d_model_per_head
    q = F.linear(q_inv, at.proj_q.weight, at.proj_q.bias).view(n, h, hc)
    k = F.linear(k_inv, at.proj_k.weight, at.proj_k.bias).view(m, h, hc)
    v = F.linear(v_eq, at.proj_v.weight, at.proj_v.bias).view(m, a, h, hc)
    s = torch.einsum('nhc,mhc->hnm', q, k) / hc ** 0.5
    hid = torch.einsum('hnm,mahc->nahc', F.softmax(s, dim=-1), v).reshape(n, a, c)
    y = F.linear(hid, al.linear.weight, al.linear.bias)
    y = F.layer_norm(y + q_inv[:, None, :], (c,), al.norm.weight, al.norm.bias, al.norm.eps)
    return _attention_output(layer.output, y)


def aten_transformer(gt, points_c, feats_c, n_ref):
    """GeometricTransformer.forward for one pair: points (T, 3), feats (T, A, Cin) -> (T, Cout)."""
    blocks = gt.transformer.blocks
    if any(b not in ('self_eq', 'cross') for b in blocks):
        raise NotImplementedError("training step: SE3ET-I blocks ('self_eq', 'cross')")
    emb_r, emb_s = _geo_embed(gt.embedding, points_c[:n_ref]), _geo_embed(gt.embedding, points_c[n_ref:])
    x = F.linear(feats_c.float(), gt.in_proj.weight, gt.in_proj.bias)
    x_inv = None
    for layer, block in zip(gt.transformer.layers, blocks):
        if block == 'self_eq':
            x = torch.cat([_rpe_layer(layer, x[:n_ref], emb_r), _rpe_layer(layer, x[n_ref:], emb_s)], 0)
            x_inv = x.amax(dim=1)
        else:
            if x_inv is None:
                x_inv = x.amax(dim=1)
            r = _cross_layer(layer, x_inv[:n_ref], x_inv[n_ref:], x[n_ref:])
            r_inv = r.amax(dim=1)
            s = _cross_layer(layer, x_inv[n_ref:], r_inv, r)     # the source attends to the UPDATED reference
            x, x_inv = torch.cat([r, s], 0), torch.cat([r_inv, s.amax(dim=1)], 0)
    return F.linear(x_inv, gt.out_proj.weight, gt.out_proj.bias)


def aten_log_optimal_transport(scores, alpha, num_iterations, row_masks, col_masks, inf=1e12):
    """LearnableLogOptimalTransport.forward (modules/sinkhorn/learnable_sinkhorn.py:13-66)."""
    b, m, n = scores.shape
    pr = torch.cat([~row_masks, torch.zeros((b, 1), dtype=torch.bool, device=scores.device)], 1)
    pc = torch.cat([~col_masks, torch.zeros((b, 1), dtype=torch.bool, device=scores.device)], 1)
    pad = pr[:, :, None] | pc[:, None, :]
    ps = torch.cat([torch.cat([scores, alpha.expand(b, m, 1)], -1), alpha.expand(b, 1, n + 1)], 1).masked_fill(pad, -inf)
    nvr, nvc = row_masks.float().sum(1), col_masks.float().sum(1)
    norm = -torch.log(nvr + nvc)
    log_mu = torch.cat([norm[:, None].expand(b, m), (torch.log(nvc) + norm)[:, None]], 1).masked_fill(pr, -inf)
    log_nu = torch.cat([norm[:, None].expand(b, n), (torch.log(nvr) + norm)[:, None]], 1).masked_fill(pc, -inf)
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(num_iterations):
        u = log_mu - torch.logsumexp(ps + v[:, None, :], dim=2)
        v = log_nu - torch.logsumexp(ps + u[:, :, None], dim=1)
    return ps + u[:, :, None] + v[:, None, :] - norm[:, None, None]


# ======================================================================================================================
# CUDA forward / ATen backward
# ======================================================================================================================
class _CoarsePath(torch.autograd.Function):
    """(model, data_dict, n_ref_c, *parameters) -> (ref_feats_c, src_feats_c, feats_f): L2-normalised superpoint
    features and fine point features.  Forward: the CUDA modules.  Backward: aten_backbone + aten_transformer."""

    @staticmethod
    def forward(ctx, model, dd, *params):
        with torch.no_grad():
            feats = torch.ones((dd['points'][0].shape[0], 1), dtype=torch.bfloat16, device=dd['points'][0].device)
            fl = model.backbone(feats, dd)
            n_ref = int(dd['lengths'][-1][0])
            pc = dd['points'][-1]
            both = model.transformer.forward_clouds(pc, fl[-1], [n_ref], [pc.shape[0] - n_ref])
            normed = F.normalize(both.float(), p=2, dim=1)
        ctx.model, ctx.dd, ctx.n_ref = model, dd, n_ref
        ctx.params = params
        return normed[:n_ref].clone(), normed[n_ref:].clone(), fl[0].float().clone()

    @staticmethod
    def backward(ctx, g_ref, g_src, g_f):
        model, dd, n_ref = ctx.model, ctx.dd, ctx.n_ref
        ac = RECOMPUTE['autocast']
        with torch.enable_grad(), torch.autocast('cuda', dtype=ac or torch.bfloat16, enabled=ac is not None):
            feats = torch.ones((dd['points'][0].shape[0], 1), dtype=torch.float32, device=g_ref.device)
            fl = aten_backbone(model.backbone, feats, dd)
            both = aten_transformer(model.transformer, dd['points'][-1], fl[-1], n_ref)
            normed = F.normalize(both.float(), p=2, dim=1)
            outs = [normed[:n_ref], normed[n_ref:], fl[0].float()]
            grads_out = [g_ref, g_src, g_f]
            live = [p for p in ctx.params if p.requires_grad]
            grads = torch.autograd.grad(outs, live, grads_out, allow_unused=True)
        it = iter(grads)
        return (None, None) + tuple(next(it) if p.requires_grad else None for p in ctx.params)


class _OptimalTransport(torch.autograd.Function):
    """Forward: se3et_log_optimal_transport (CUDA).  Backward: the log-domain Sinkhorn iterations in ATen."""

    @staticmethod
    def forward(ctx, scores, alpha, num_iterations, row_masks, col_masks):
        ctx.save_for_backward(scores, alpha, row_masks, col_masks)
        ctx.iters = num_iterations
        with torch.no_grad():
            return log_optimal_transport(scores, alpha, num_iterations, row_masks, col_masks)

    @staticmethod
    def backward(ctx, g):
        scores, alpha, rm, cm = ctx.saved_tensors
        with torch.enable_grad():
            s = scores.detach().requires_grad_(True)
            a = alpha.detach().requires_grad_(True)
            out = aten_log_optimal_transport(s, a, ctx.iters, rm, cm)
            # masked entries carry -inf-like values with zero upstream gradient
            gs, ga = torch.autograd.grad(out, (s, a), torch.where(out > -1e11, g, torch.zeros_like(g)))
        return gs, ga, None, None, None


# ======================================================================================================================
# ground truth and losses
# ======================================================================================================================
def apply_transform(points, transform):
    return points @ transform[:3, :3].t() + transform[:3, 3]


def _sq_dist(x, y):
    """pairwise_distance (modules/ops/pairwise_distance.py:4-31), not normalised."""
    xy = x @ y.transpose(-1, -2)
    d = (x ** 2).sum(-1).unsqueeze(-1) - 2 * xy + (y ** 2).sum(-1).unsqueeze(-2)
    return d.clamp(min=0.0)


@torch.no_grad()
def get_node_correspondences(ref_nodes, src_nodes, ref_knn_points, src_knn_points, transform, pos_radius, ref_masks,
                             src_masks, ref_knn_masks, src_knn_masks):
    """modules/registration/matching.py:231-315 -> corr_indices (C, 2), corr_overlaps (C,)."""
    src_nodes = apply_transform(src_nodes, transform)
    src_knn_points = apply_transform(src_knn_points, transform)
    node_mask_mat = ref_masks[:, None] & src_masks[None, :]
    ref_d = torch.linalg.norm(ref_knn_points - ref_nodes[:, None], dim=-1).masked_fill(~ref_knn_masks, 0.0)
    src_d = torch.linalg.norm(src_knn_points - src_nodes[:, None], dim=-1).masked_fill(~src_knn_masks, 0.0)
    dist_mat = torch.sqrt(_sq_dist(ref_nodes, src_nodes))
    inter = (ref_d.max(1)[0][:, None] + src_d.max(1)[0][None, :] + pos_radius - dist_mat > 0) & node_mask_mat
    sel_r, sel_s = torch.nonzero(inter, as_tuple=True)
    rkm, skm = ref_knn_masks[sel_r], src_knn_masks[sel_s]
    d = _sq_dist(ref_knn_points[sel_r], src_knn_points[sel_s]).masked_fill(~(rkm[:, :, None] & skm[:, None, :]), 1e12)
    ov = d < pos_radius ** 2
    ref_ov = torch.count_nonzero(ov.sum(-1), dim=-1).float() / rkm.sum(-1).float()
    src_ov = torch.count_nonzero(ov.sum(-2), dim=-1).float() / skm.sum(-1).float()
    overlaps = (ref_ov + src_ov) / 2
    keep = overlaps > 0
    return torch.stack([sel_r[keep], sel_s[keep]], dim=1), overlaps[keep]


def weighted_circle_loss(pos_masks, neg_masks, feat_dists, pos_margin, neg_margin, pos_optimal, neg_optimal, log_scale,
                         pos_scales=None, neg_scales=None):
    """modules/loss/circle_loss.py:44-87."""
    row_masks = ((pos_masks.sum(-1) > 0) & (neg_masks.sum(-1) > 0)).detach()
    col_masks = ((pos_masks.sum(-2) > 0) & (neg_masks.sum(-2) > 0)).detach()
    pos_w = torch.clamp(feat_dists - 1e5 * (~pos_masks).float() - pos_optimal, min=0.0)
    if pos_scales is not None:
        pos_w = pos_w * pos_scales
    pos_w = pos_w.detach()
    neg_w = torch.clamp(neg_optimal - (feat_dists + 1e5 * (~neg_masks).float()), min=0.0)
    if neg_scales is not None:
        neg_w = neg_w * neg_scales
    neg_w = neg_w.detach()
    lp = log_scale * (feat_dists - pos_margin) * pos_w
    ln = log_scale * (neg_margin - feat_dists) * neg_w
    loss_row = F.softplus(torch.logsumexp(lp, dim=-1) + torch.logsumexp(ln, dim=-1)) / log_scale
    loss_col = F.softplus(torch.logsumexp(lp, dim=-2) + torch.logsumexp(ln, dim=-2)) / log_scale
    return (loss_row[row_masks].mean() + loss_col[col_masks].mean()) / 2


class LossCfg:
    """config.py:220-236 of every SE3ET 3DMatch experiment."""
    positive_margin, negative_margin, positive_optimal, negative_optimal, log_scale = 0.1, 1.4, 0.1, 1.4, 24
    positive_overlap, positive_radius = 0.1, 0.05
    weight_coarse_loss, weight_fine_loss = 1.0, 1.0
    ground_truth_matching_radius, num_targets, overlap_threshold = 0.05, 128, 0.1


def coarse_matching_loss(ref_feats, src_feats, gt_indices, gt_overlaps, cfg=LossCfg):
    """CoarseMatchingLoss.forward (experiments/se3eti.3dmatch/loss.py:15-45)."""
    feat_dists = torch.sqrt((2.0 - 2.0 * ref_feats @ src_feats.t()).clamp(min=0.0))
    overlaps = torch.zeros_like(feat_dists)
    overlaps[gt_indices[:, 0], gt_indices[:, 1]] = gt_overlaps
    pos_masks = overlaps > cfg.positive_overlap
    neg_masks = overlaps == 0
    pos_scales = torch.sqrt(overlaps * pos_masks.float())
    return weighted_circle_loss(pos_masks, neg_masks, feat_dists, cfg.positive_margin, cfg.negative_margin,
                                cfg.positive_optimal, cfg.negative_optimal, cfg.log_scale, pos_scales)


def fine_matching_loss(matching_scores, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, transform,
                       cfg=LossCfg):
    """FineMatchingLoss.forward (experiments/se3eti.3dmatch/loss.py:48-77)."""
    d = _sq_dist(ref_knn_points, apply_transform(src_knn_points, transform))
    gt = (d < cfg.positive_radius ** 2) & (ref_knn_masks[:, :, None] & src_knn_masks[:, None, :])
    slack_row = (gt.sum(2) == 0) & ref_knn_masks
    slack_col = (gt.sum(1) == 0) & src_knn_masks
    labels = torch.zeros_like(matching_scores, dtype=torch.bool)
    labels[:, :-1, :-1] = gt
    labels[:, :-1, -1] = slack_row
    labels[:, -1, :-1] = slack_col
    return -matching_scores[labels].mean()


# ======================================================================================================================
# the step
# ======================================================================================================================
def trainable_parameters(model):
    return [p for p in model.parameters() if p.requires_grad]


def allreduce_gradients(params, world_size):
    """Average the gradients over the ranks with ONE all-reduce of a flattened bucket (what DistributedDataParallel
    does bucket by bucket in the reference, engine/base_trainer.py:181-189)."""
    if world_size <= 1:
        return 0
    import torch.distributed as dist
    grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size)
    o = 0
    for p, g in zip(params, grads):
        n = g.numel()
        p.grad = flat[o:o + n].view_as(g).clone()
        o += n
    return flat.numel() * flat.element_size()


def training_step(model, ref_points, src_points, transform, optimizer=None, world_size=1, rng=None, cfg=LossCfg):
    """One iteration on one pair (host arrays in): CUDA forward, losses, ATen backward, gradient exchange, optimizer
    step.  Returns a dict of python floats (loss, c_loss, f_loss) and the number of gradient bytes exchanged."""
    dev = next(model.parameters()).device
    mcfg = model.cfg
    b = mcfg.backbone
    pts = torch.from_numpy(np.concatenate([ref_points, src_points]).astype(np.float32)).to(dev)
    lens = torch.tensor([len(ref_points), len(src_points)], dtype=torch.int64, device=dev)
    tr = torch.as_tensor(np.asarray(transform, dtype=np.float32), device=dev)
    with torch.no_grad():
        dd = precompute_data_stack_mode(pts, lens, b.num_stages, b.init_voxel_size, b.init_radius, mcfg.neighbor_limits)
    params = trainable_parameters(model)
    ref_c, src_c, feats_f = _CoarsePath.apply(model, dd, *params)

    with torch.no_grad():
        n_ref_c, n_ref_f = int(dd['lengths'][-1][0]), int(dd['lengths'][1][0])
        pc, pf = dd['points'][-1], dd['points'][1]
        k = mcfg.model.num_points_in_patch
        _, node_masks, knn, knn_masks = point_to_node_partition_stacked(pf, dd['lengths'][1], pc, dd['lengths'][-1], k)
        rk, sk = knn[:n_ref_c], knn[n_ref_c:]
        rkm, skm = knn_masks[:n_ref_c], knn_masks[n_ref_c:]
        ref_f_pad = torch.cat([pf[:n_ref_f], torch.zeros_like(pf[:1])])
        src_f_pad = torch.cat([pf[n_ref_f:], torch.zeros_like(pf[:1])])
        gt_idx, gt_ov = get_node_correspondences(pc[:n_ref_c], pc[n_ref_c:], ref_f_pad[rk], src_f_pad[sk], tr,
                                                 cfg.ground_truth_matching_radius, node_masks[:n_ref_c],
                                                 node_masks[n_ref_c:], rkm, skm)
        # SuperPointTargetGenerator: up to num_targets ground-truth patch pairs above the overlap threshold
        sel = torch.nonzero(gt_ov > cfg.overlap_threshold, as_tuple=True)[0]
        if sel.numel() > cfg.num_targets:
            rng = rng or np.random.default_rng(0)
            pick = rng.choice(sel.numel(), cfg.num_targets, replace=False)
            sel = sel[torch.from_numpy(pick).to(dev)]
        t_ref, t_src = gt_idx[sel, 0], gt_idx[sel, 1]
    if gt_idx.shape[0] == 0 or sel.numel() == 0:
        raise RuntimeError("training_step: the pair has no ground-truth superpoint correspondence")

    c_loss = coarse_matching_loss(ref_c, src_c, gt_idx, gt_ov, cfg)
    # fine stage on the ground-truth patch pairs (model.py:175-205 of the reference, training branch)
    ref_ff = torch.cat([feats_f[:n_ref_f], torch.zeros_like(feats_f[:1])])
    src_ff = torch.cat([feats_f[n_ref_f:], torch.zeros_like(feats_f[:1])])
    scores = torch.einsum('bnd,bmd->bnm', gather_rows(ref_ff, rk[t_ref]), gather_rows(src_ff, sk[t_src])) / \
        feats_f.shape[1] ** 0.5
    ms = _OptimalTransport.apply(scores, model.optimal_transport.alpha, model.optimal_transport.num_iterations,
                                 rkm[t_ref], skm[t_src])
    f_loss = fine_matching_loss(ms, ref_f_pad[rk[t_ref]], src_f_pad[sk[t_src]], rkm[t_ref], skm[t_src], tr, cfg)
    loss = cfg.weight_coarse_loss * c_loss + cfg.weight_fine_loss * f_loss

    for p in params:
        p.grad = None
    loss.backward()
    nbytes = allreduce_gradients(params, world_size)
    if optimizer is not None:
        optimizer.step()
    return {'loss': float(loss.detach()), 'c_loss': float(c_loss.detach()), 'f_loss': float(f_loss.detach()),
            'grad_bytes': nbytes}
