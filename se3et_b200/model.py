"""Hot-path model wiring: mirror of experiments/<variant>/{config,model}.py up to and including coarse matching.

    cfg = make_cfg('se3eti.3dmatch')            # experiments/se3eti.3dmatch/config.py:60-239 (the fields the path reads)
    model = create_model(cfg).cuda().eval()     # experiments/se3eti.3dmatch/model.py:231-233
    out = model(data_dict)                      # one pair, reference data_dict (model.py:80-172)
    out = model.forward_pairs(clouds)           # P pairs in one launch sequence, host arrays in / device tensors out

state_dict keys are the reference's (`backbone.*`, `transformer.*`), so a reference checkpoint loads with
strict=False (the fine-matching / optimal-transport modules downstream of SuperPointMatching are out of scope).
"""
import numpy as np
import torch
import torch.nn as nn

from . import _lib
from .modules.e2pn import E2PN
from .modules.transformer import GeometricTransformer, SuperPointMatching
from .ops import transformer_ops as T
from .modules.registration import LocalGlobalRegistration
from .modules.sinkhorn import LearnableLogOptimalTransport
from .ops.gemm import bmm_bf16
from .ops.partition_ops import point_to_node_partition_stacked
from .precompute import precompute_data_stack_mode


class Cfg(dict):
    """Attribute dict (stand-in for easydict, which the reference configs use)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


_VARIANTS = {
    # name: (init_dim, group_norm, backbone_out, hidden, tr_out, stages, voxel, base_radius, sigma_d, limits)
    'se3eti.3dmatch': (64, 32, 256, 256, 256, 4, 0.025, 2.5, 0.2, [38, 36, 36, 38]),
    'se3eti2.3dmatch': (32, 16, 128, 128, 128, 4, 0.025, 2.5, 0.2, [38, 36, 36, 38]),
    'se3eti.kitti': (64, 32, 256, 128, 256, 5, 0.3, 4.25, 4.8, [40, 40, 40, 40, 40]),
    'se3ete.3dmatch': (64, 32, 256, 256, 256, 4, 0.025, 2.5, 0.2, [38, 36, 36, 38]),
    'se3ete2.3dmatch': (32, 16, 128, 128, 128, 4, 0.025, 2.5, 0.2, [38, 36, 36, 38]),
}
_BLOCKS_I = ['self_eq', 'cross', 'self_eq', 'cross', 'self_eq', 'cross']
# experiments/se3ete.3dmatch/config.py:194
_BLOCKS_E = ['self_eq', 'cross_a_soft', 'self_eq', 'cross_r_soft', 'self', 'cross', 'self', 'cross', 'self', 'cross']


def make_cfg(variant='se3eti.3dmatch'):
    if variant not in _VARIANTS:
        raise NotImplementedError("variant %r: the CUDA path covers %s" % (variant, sorted(_VARIANTS)))
    d, g, bo, hid, to, stages, voxel, base_r, sigma_d, limits = _VARIANTS[variant]
    c = Cfg(variant=variant)
    is_e = variant.startswith('se3ete')
    c.backbone = Cfg(num_stages=stages, init_voxel_size=voxel, kernel_size=15, base_radius=base_r, base_sigma=2.0,
                     init_radius=base_r * voxel, init_sigma=2.0 * voxel, group_norm=g, input_dim=1, init_dim=d,
                     output_dim=bo)
    c.epn = Cfg(kanchor=6, quotient_factor=4, num_kernel_points=15, non_sep_conv=True, equiv_mode_kp=True,
                fixed_kernel_points='center', rot_by_permute=True, ignore_steer_constraint=False, epn_kernel=False,
                att_pooling=False, att_permute=False, dual_feature=False, gather_by_idxing=False,
                use_batch_norm=True, batch_norm_momentum=0.99, KP_extent=1.0, KP_influence='linear',
                aggregation_mode='sum')
    c.geotransformer = Cfg(input_dim=d * (16 if stages == 4 else 32), hidden_dim=hid, output_dim=to, num_heads=4,
                           blocks=list(_BLOCKS_E if is_e else _BLOCKS_I), sigma_d=sigma_d,
                           sigma_a=15, angle_k=3, supervise_rotation=False, reduction_a='max', align_mode='0',
                           alternative_impl=False, n_level_equiv=2 if is_e else 0,
                           # SE3ET-E's model.py does not pass attn_r_positive*: the module defaults apply
                           attn_r_positive='sq' if is_e else 'softplus',
                           attn_r_positive_rot_supervise='sigmoid' if is_e else 'minus')
    # config.py:177-178 (3DMatch) / se3eti.kitti/config.py:180-181
    c.model = Cfg(num_points_in_patch=128 if stages == 5 else 64, num_sinkhorn_iterations=100)
    c.coarse_matching = Cfg(num_targets=128, overlap_threshold=0.1, num_correspondences=256, dual_normalization=True)
    # config.py:208-217 (3DMatch) / se3eti.kitti/config.py:212-221
    c.fine_matching = Cfg(topk=2 if stages == 5 else 3, acceptance_radius=0.6 if stages == 5 else 0.1, mutual=True,
                          confidence_threshold=0.05, use_dustbin=False, use_global_score=False,
                          correspondence_threshold=3, correspondence_limit=None, num_refinement_steps=5)
    c.neighbor_limits = list(limits)  # demo.py:52 for 3DMatch; KITTI limits are calibrated per dataset (data.py:212-252)
    return c


class GeoTransformer(nn.Module):
    """experiments/se3eti.3dmatch/model.py:20-172 (KPFCNN encoder -> conditional transformer -> coarse matching)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        b, g = cfg.backbone, cfg.geotransformer
        self.backbone = E2PN(b.input_dim, b.output_dim, b.init_dim, b.init_radius, b.init_sigma, b.group_norm, cfg.epn,
                             num_stages=b.num_stages)
        self.transformer = GeometricTransformer(
            g.input_dim, g.output_dim, g.hidden_dim, g.num_heads, g.blocks, g.sigma_d, g.sigma_a, g.angle_k,
            supervise_rotation=g.supervise_rotation, reduction_a=g.reduction_a, na=cfg.epn.kanchor,
            attn_r_positive=g.attn_r_positive, attn_r_positive_rot_supervise=g.attn_r_positive_rot_supervise,
            align_mode=g.align_mode, alternative_impl=g.alternative_impl, n_level_equiv=g.n_level_equiv)
        self.coarse_matching = SuperPointMatching(cfg.coarse_matching.num_correspondences,
                                                  cfg.coarse_matching.dual_normalization)
        self.optimal_transport = LearnableLogOptimalTransport(cfg.model.num_sinkhorn_iterations)  # model.py:76
        f = cfg.fine_matching
        self.fine_matching = LocalGlobalRegistration(                                              # model.py:63-73
            f.topk, f.acceptance_radius, mutual=f.mutual, confidence_threshold=f.confidence_threshold,
            use_dustbin=f.use_dustbin, use_global_score=f.use_global_score,
            correspondence_threshold=f.correspondence_threshold, correspondence_limit=f.correspondence_limit,
            num_refinement_steps=f.num_refinement_steps)

    @torch.no_grad()
    def fine_matching_scores(self, out):
        """Steps 7.2-8 of the reference forward (experiments/se3eti.3dmatch/model.py:184-205) on the output of
        forward(): patch features of the selected superpoint pairs, (P, K, K) scores / sqrt(C) as one batched tcgen05
        GEMM, log-domain optimal transport.  -> matching_scores (P, K + 1, K + 1) fp32 (also stored in `out`)."""
        ri, si = out['ref_node_corr_indices'], out['src_node_corr_indices']
        feats = []
        for f, knn, idx in ((out['ref_feats_f'], out['ref_node_knn_indices'], ri),
                            (out['src_feats_f'], out['src_node_knn_indices'], si)):
            f = f.to(torch.bfloat16)
            padded = torch.cat([f, torch.zeros_like(f[:1])], dim=0)                 # the shadow row (model.py:191-192)
            k_idx = knn.index_select(0, idx)                                       # (P, K)
            feats.append(padded.index_select(0, k_idx.reshape(-1)).view(k_idx.shape[0], k_idx.shape[1], -1).contiguous())
        c = feats[0].shape[-1]
        scores, _ = bmm_bf16(feats[0], feats[1], alpha=1.0 / c ** 0.5)
        ms = self.optimal_transport(scores, out['ref_node_knn_masks'].index_select(0, ri),
                                    out['src_node_knn_masks'].index_select(0, si))
        out['matching_scores'] = ms
        return ms

    @torch.no_grad()
    def register(self, out, data_dict):
        """Steps 7.2-9 of the reference forward (experiments/se3eti.3dmatch/model.py:175-224) after forward(): optimal
        transport on the patch pairs, LocalGlobalRegistration.  Adds matching_scores, ref_corr_points, src_corr_points,
        corr_scores and estimated_transform (4, 4) to `out`."""
        ms = self.fine_matching_scores(out)
        points_f = data_dict['points'][1]
        ref_length_f = int(data_dict['lengths'][1][0])
        ri, si = out['ref_node_corr_indices'], out['src_node_corr_indices']
        knn_pts = []
        for pts, knn, idx in ((points_f[:ref_length_f], out['ref_node_knn_indices'], ri),
                              (points_f[ref_length_f:], out['src_node_knn_indices'], si)):
            padded = torch.cat([pts, torch.zeros_like(pts[:1])], dim=0)            # model.py:117-118 (shadow point)
            k_idx = knn.index_select(0, idx)
            knn_pts.append(padded.index_select(0, k_idx.reshape(-1)).view(k_idx.shape[0], k_idx.shape[1], 3))
        rp, sp, sc, t = self.fine_matching(knn_pts[0], knn_pts[1], out['ref_node_knn_masks'].index_select(0, ri),
                                           out['src_node_knn_masks'].index_select(0, si), ms, out['node_corr_scores'])
        out['ref_corr_points'], out['src_corr_points'], out['corr_scores'], out['estimated_transform'] = rp, sp, sc, t
        return out

    @torch.no_grad()
    def register_stacked(self, res):
        """Fine stage for the P pairs of forward_stacked(): patch gather + batched score GEMM + optimal transport +
        LocalGlobalRegistration for all P x num_correspondences patch pairs in one launch sequence.
        Adds 'estimated_transforms' (P, 4, 4), 'ref_corr_points' / 'src_corr_points' / 'corr_scores' (stacked) and
        'corr_offsets' (P + 1,) to `res`."""
        dd = res['data_dict']
        dev = res['feats_f'].device
        num_pairs = len(res['ref_sizes'])
        k = self.cfg.model.num_points_in_patch
        points_f, feats_f = dd['points'][1], res['feats_f']
        len_f = dd['lengths'][1].to(dev)
        len_c = dd['lengths'][-1].to(dev)
        start_f = torch.cumsum(len_f, 0) - len_f                    # first fine point of every cloud (backbone order)
        start_c = torch.cumsum(len_c, 0) - len_c
        cnt = res['num_corr'].to(torch.int64)                       # correspondences per pair
        ncor = res['ref_node_corr_indices'].shape[1]
        slot = torch.arange(ncor, device=dev)[None, :].expand(num_pairs, -1)
        live = slot < cnt[:, None]
        pair_id = torch.arange(num_pairs, device=dev)[:, None].expand(-1, ncor)[live]
        knn, knn_masks = res['node_knn_indices'], res['node_knn_masks']       # cloud-local point ids, backbone order
        feats = feats_f.to(torch.bfloat16)
        zf = torch.zeros_like(feats[:1])
        zp = torch.zeros_like(points_f[:1])
        g_feats, g_pts, g_masks = [], [], []
        for side, corr in ((0, res['ref_node_corr_indices']), (1, res['src_node_corr_indices'])):
            cloud = 2 * pair_id + side
            node = start_c[cloud] + corr[live]                      # node row in backbone order
            k_idx = knn.index_select(0, node)                       # (B, K) cloud-local, shadow = cloud length
            m = knn_masks.index_select(0, node)
            rows = torch.where(m, k_idx + start_f[cloud][:, None], torch.full_like(k_idx, feats.shape[0]))
            g_feats.append(torch.cat([feats, zf]).index_select(0, rows.reshape(-1)).view(rows.shape[0], k, -1).contiguous())
            g_pts.append(torch.cat([points_f, zp]).index_select(0, rows.reshape(-1)).view(rows.shape[0], k, 3))
            g_masks.append(m)
        c = g_feats[0].shape[-1]
        scores, _ = bmm_bf16(g_feats[0], g_feats[1], alpha=1.0 / c ** 0.5)
        ms = self.optimal_transport(scores, g_masks[0], g_masks[1])
        patch_off = torch.cat([cnt.new_zeros(1), torch.cumsum(cnt, 0)])
        rp, sp, sc, coff, tr = self.fine_matching.forward_pairs(g_pts[0], g_pts[1], g_masks[0], g_masks[1], ms, patch_off)
        res.update(matching_scores=ms, ref_corr_points=rp, src_corr_points=sp, corr_scores=sc, corr_offsets=coff,
                   estimated_transforms=tr)
        return res

    # ---- one pair, reference data_dict -------------------------------------------------------------------------
    @torch.no_grad()
    def forward(self, data_dict, ref_node_masks=None, src_node_masks=None):
        out = {}
        feats = data_dict['features']
        ref_length_c = int(data_dict['lengths'][-1][0])
        ref_length_f = int(data_dict['lengths'][1][0])
        points_c = data_dict['points'][-1]
        feats_list = self.backbone(feats, data_dict)
        feats_c, feats_f = feats_list[-1], feats_list[0]
        n_c = points_c.shape[0]
        ref_c, src_c = self.transformer.forward_clouds(points_c, feats_c, [ref_length_c], [n_c - ref_length_c]).split(
            [ref_length_c, n_c - ref_length_c])
        ref_n = T.l2_normalize_rows(ref_c.contiguous())
        src_n = T.l2_normalize_rows(src_c.contiguous())
        out['ref_points_c'], out['src_points_c'] = points_c[:ref_length_c], points_c[ref_length_c:]
        out['ref_feats_c'], out['src_feats_c'] = ref_n, src_n
        out['ref_feats_f'], out['src_feats_f'] = feats_f[:ref_length_f], feats_f[ref_length_f:]
        # point-to-node partition of the fine level (model.py:109-119 of the reference): node masks for the coarse
        # matching, patch indices / masks for the fine stage
        points_f = data_dict['points'][1]
        lf = data_dict['lengths'][1].to(points_f.device)
        lc = data_dict['lengths'][-1].to(points_f.device)
        _, node_masks, knn, knn_masks = point_to_node_partition_stacked(points_f, lf, points_c, lc,
                                                                         self.cfg.model.num_points_in_patch)
        if ref_node_masks is None:
            ref_node_masks = node_masks[:ref_length_c]
        if src_node_masks is None:
            src_node_masks = node_masks[ref_length_c:]
        out['ref_node_masks'], out['src_node_masks'] = ref_node_masks, src_node_masks
        out['ref_node_knn_indices'], out['src_node_knn_indices'] = knn[:ref_length_c], knn[ref_length_c:]
        out['ref_node_knn_masks'], out['src_node_knn_masks'] = knn_masks[:ref_length_c], knn_masks[ref_length_c:]
        ri, si, sc = self.coarse_matching(ref_n, src_n, ref_node_masks, src_node_masks)
        out['ref_node_corr_indices'], out['src_node_corr_indices'], out['node_corr_scores'] = ri, si, sc
        return out

    # ---- P pairs, one launch sequence ---------------------------------------------------------------------------
    @torch.no_grad()
    def forward_stacked(self, points, lengths):
        """points fp32 (N, 3) on the GPU, clouds stacked [ref_0, src_0, ref_1, src_1, ...]; lengths int64 (2P,) on the
        HOST. Returns dict with per-pair correspondences (P, k) and the flat coarse features."""
        _lib.require_cuda(points)
        cfg = self.cfg
        num_pairs = lengths.shape[0] // 2
        dev = points.device
        b = cfg.backbone
        dd = precompute_data_stack_mode(points, lengths.to(dev), b.num_stages, b.init_voxel_size, b.init_radius,
                                        cfg.neighbor_limits, backbone_only=num_pairs > 1)
        feats = torch.ones((points.shape[0], 1), dtype=torch.bfloat16, device=dev)
        feats_list = self.backbone(feats, dd)
        feats_c, feats_f = feats_list[-1], feats_list[0]
        len_c = dd['lengths'][-1].cpu().numpy()  # already synchronised by the subsampling that produced it
        if num_pairs > 1 and int(len_c.max(initial=0)) > 2000:
            # the reference keeps at most 2000 superpoints per cloud (utils/data.py:34-43, applied in pair mode by
            # precompute_data_stack_mode); a stacked launch would silently differ from the pair run alone
            raise RuntimeError("forward_stacked: a cloud has %d superpoints (> 2000); run it through forward()" %
                               int(len_c.max()))
        ref_sizes, src_sizes = len_c[0::2], len_c[1::2]
        points_c = dd['points'][-1]
        if num_pairs > 1:
            # transformer order: all reference clouds, then all source clouds
            starts = np.concatenate([[0], np.cumsum(len_c)])[:-1]
            order = np.concatenate([np.arange(starts[i], starts[i] + len_c[i]) for i in
                                    list(range(0, 2 * num_pairs, 2)) + list(range(1, 2 * num_pairs, 2))])
            order_t = torch.from_numpy(order).to(dev, non_blocking=True)
            points_c = points_c.index_select(0, order_t)
            feats_c = feats_c.index_select(0, order_t)
        tr = int(ref_sizes.sum())
        # point-to-node partition of every cloud (fine level = stage 2, nodes = last stage), backbone order
        _, node_masks, knn, knn_masks = point_to_node_partition_stacked(
            dd['points'][1], dd['lengths'][1], dd['points'][-1], dd['lengths'][-1], cfg.model.num_points_in_patch)
        masks_t = node_masks if num_pairs == 1 else node_masks.index_select(0, order_t)
        both = self.transformer.forward_clouds(points_c, feats_c, ref_sizes.tolist(), src_sizes.tolist())
        normed = T.l2_normalize_rows(both)
        ri, si, sc, cnt = self.coarse_matching.forward_pairs(normed[:tr], normed[tr:], ref_sizes, src_sizes,
                                                             masks_t[:tr].to(torch.uint8), masks_t[tr:].to(torch.uint8))
        return {
            'node_masks': node_masks, 'node_knn_indices': knn, 'node_knn_masks': knn_masks,
            'ref_node_corr_indices': ri, 'src_node_corr_indices': si, 'node_corr_scores': sc, 'num_corr': cnt,
            'ref_feats_c': normed[:tr], 'src_feats_c': normed[tr:], 'ref_sizes': ref_sizes, 'src_sizes': src_sizes,
            'feats_f': feats_f, 'points_c': points_c, 'data_dict': dd,
        }

    def _stream_pool(self, dev, n):
        if not hasattr(self, '_streams') or len(self._streams) < n:
            self._streams = [torch.cuda.Stream(device=dev) for _ in range(n)]
        return self._streams

    @torch.no_grad()
    def forward_stacked_concurrent(self, inputs, num_streams=2):
        """Runs several launch sequences (`inputs` = [(points, lengths), ...] as for forward_stacked) concurrently,
        one host thread + CUDA stream per sequence slot: pairs are independent, and a second sequence fills the SMs
        while the first sits in a latency-bound kernel or a host-side size read-back.  Returns the list of results."""
        dev = inputs[0][0].device
        if num_streams <= 1 or len(inputs) <= 1:
            return [self.forward_stacked(p, l) for p, l in inputs]
        return _run_concurrent(lambda i: self.forward_stacked(*inputs[i]), len(inputs), num_streams,
                               self._stream_pool(dev, num_streams), dev)

    @torch.no_grad()
    def forward_pairs_concurrent(self, groups, num_streams=2, pinned=None):
        """forward_pairs for several groups of pairs at once (host arrays in, host results out), `num_streams`
        groups in flight.  pinned: optional list of num_streams pinned staging buffers."""
        dev = next(self.parameters()).device
        if num_streams <= 1 or len(groups) <= 1:
            return [self.forward_pairs(g, pinned=None if pinned is None else pinned[0]) for g in groups]
        return _run_concurrent(
            lambda i: self.forward_pairs(groups[i], pinned=None if pinned is None else pinned[i % num_streams]),
            len(groups), num_streams, self._stream_pool(dev, num_streams), dev)

    @torch.no_grad()
    def forward_pairs(self, clouds, pinned=None):
        """Public end-to-end entry: clouds = [(ref (n,3) float32 ndarray, src (m,3) float32 ndarray), ...] on the HOST.
        Copies the stacked points to the GPU (from pinned memory), runs the whole hot path and returns the
        correspondences on the HOST: list of (ref_idx, src_idx, scores) numpy arrays."""
        lens = np.array([len(c) for pair in clouds for c in pair], dtype=np.int64)
        total = int(lens.sum())
        if pinned is None or pinned.shape[0] < total:
            pinned = torch.empty((total, 3), dtype=torch.float32).pin_memory()
        o = 0
        host = pinned.numpy()
        for pair in clouds:
            for c in pair:
                host[o:o + len(c)] = c
                o += len(c)
        dev = next(self.parameters()).device
        points = pinned[:total].to(dev, non_blocking=True)
        out = self.forward_stacked(points, torch.from_numpy(lens))
        ri = out['ref_node_corr_indices'].cpu().numpy()
        si = out['src_node_corr_indices'].cpu().numpy()
        sc = out['node_corr_scores'].cpu().numpy()
        cnt = out['num_corr'].cpu().numpy()
        return [(ri[p, :cnt[p]], si[p, :cnt[p]], sc[p, :cnt[p]]) for p in range(len(clouds))]


def _run_concurrent(fn, count, num_streams, streams, dev):
    """fn(i) for i in range(count), slot s runs i = s, s + num_streams, ... on its own host thread and CUDA stream."""
    import threading
    results = [None] * count
    errors = []
    start = torch.cuda.Event()
    start.record(torch.cuda.current_stream(dev))

    def worker(slot):
        try:
            torch.cuda.set_device(dev)
            st = streams[slot]
            st.wait_event(start)
            with torch.cuda.stream(st), torch.no_grad():
                for i in range(slot, count, num_streams):
                    results[i] = fn(i)
        except Exception as e:  # surfaced in the caller
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(s,)) for s in range(num_streams)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    cur = torch.cuda.current_stream(dev)
    for st in streams[:num_streams]:
        cur.wait_stream(st)
    return results


def create_model(config):
    return GeoTransformer(config)
