"""Drop-ins for geotransformer/modules/registration/procrustes.py (weighted_procrustes, WeightedProcrustes) and
geotransformer/modules/geotransformer/local_global_registration.py (LocalGlobalRegistration) on the CUDA path: same
names, constructor arguments and forward signatures; `forward_pairs` registers all pairs of a launch sequence at once."""
import ctypes

import torch
import torch.nn as nn

from .. import _lib


def weighted_procrustes(src_points, ref_points, weights=None, weight_thresh=0.0, eps=1e-5, return_transform=False):
    """procrustes.py:6-73: rigid transform from src onto ref, (N, 3) or (B, N, 3) fp32 CUDA tensors."""
    _lib.require_cuda(src_points, ref_points, weights)
    squeeze = src_points.dim() == 2
    src = src_points.float().reshape(-1, src_points.shape[-2], 3).contiguous()
    ref = ref_points.float().reshape(-1, ref_points.shape[-2], 3).contiguous()
    b, n = src.shape[0], src.shape[1]
    w = None if weights is None else weights.float().reshape(b, n).contiguous()
    out = torch.empty((b, 4, 4), dtype=torch.float32, device=src.device)
    _lib.check(_lib.lib().se3et_weighted_procrustes(_lib.ptr(src), _lib.ptr(ref), _lib.ptr(w), _lib.i64(b), _lib.i64(n),
                                                    _lib.f32(weight_thresh), _lib.f32(eps), _lib.ptr(out),
                                                    _lib.stream_ptr()), "weighted_procrustes")
    if return_transform:
        return out[0] if squeeze else out
    R, t = out[:, :3, :3], out[:, :3, 3]
    return (R[0], t[0]) if squeeze else (R, t)


class WeightedProcrustes(nn.Module):
    """procrustes.py:76-92."""

    def __init__(self, weight_thresh=0.0, eps=1e-5, return_transform=False):
        super().__init__()
        self.weight_thresh, self.eps, self.return_transform = weight_thresh, eps, return_transform

    def forward(self, src_points, tgt_points, weights=None):
        return weighted_procrustes(src_points, tgt_points, weights=weights, weight_thresh=self.weight_thresh,
                                   eps=self.eps, return_transform=self.return_transform)


class LocalGlobalRegistration(nn.Module):
    """local_global_registration.py:12-235 (mutual matching without dustbin / global scores / correspondence limit:
    the configuration every SE3ET experiment uses, config.py:208-217)."""

    def __init__(self, k, acceptance_radius, mutual=True, confidence_threshold=0.05, use_dustbin=False,
                 use_global_score=False, correspondence_threshold=3, correspondence_limit=None, num_refinement_steps=5):
        super().__init__()
        if not mutual or use_dustbin or use_global_score or correspondence_limit is not None:
            raise NotImplementedError("se3et_b200.LocalGlobalRegistration: mutual matching without dustbin, global "
                                      "scores or a correspondence limit (the SE3ET configuration)")
        self.k, self.acceptance_radius, self.mutual = k, acceptance_radius, mutual
        self.confidence_threshold, self.use_dustbin, self.use_global_score = confidence_threshold, use_dustbin, use_global_score
        self.correspondence_threshold, self.correspondence_limit = correspondence_threshold, correspondence_limit
        self.num_refinement_steps = num_refinement_steps
        self.procrustes = WeightedProcrustes(return_transform=True)

    @torch.no_grad()
    def forward_pairs(self, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat, patch_offsets):
        """All pairs of a launch sequence: (B, K, 3) x 2, (B, K) masks, score_mat (B, K or K + 1, K or K + 1) log scores
        (a trailing dustbin row / column is ignored), patch_offsets int64 (P + 1,) on the GPU.
        -> ref_corr_points, src_corr_points (C, 3), corr_scores (C,), corr_offsets int64 (P + 1,) on the GPU (pair i owns
        [corr_offsets[i], corr_offsets[i + 1])), transforms (P, 4, 4).  No host sync except the final size read."""
        _lib.require_cuda(ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat, patch_offsets)
        b, k = ref_knn_masks.shape
        assert score_mat.dim() == 3 and score_mat.shape[0] == b and score_mat.shape[1] == score_mat.shape[2] >= k
        dev = score_mat.device
        score_mat = score_mat.float().contiguous()
        rp, sp = ref_knn_points.float().contiguous(), src_knn_points.float().contiguous()
        rm, sm = ref_knn_masks.to(torch.uint8).contiguous(), src_knn_masks.to(torch.uint8).contiguous()
        patch_offsets = patch_offsets.to(torch.int64).contiguous()
        num_pairs = patch_offsets.numel() - 1
        L = _lib.lib()
        nbytes = ctypes.c_size_t(0)
        _lib.check(L.se3et_lgr_workspace_bytes(_lib.i64(b), _lib.i64(k), _lib.i64(self.k), ctypes.byref(nbytes)),
                   "lgr_workspace_bytes")
        ws = torch.empty((nbytes.value + 256,), dtype=torch.uint8, device=dev)
        off = (-ws.data_ptr()) % 256
        wptr, wsize = ctypes.c_void_p(ws.data_ptr() + off), ctypes.c_size_t(ws.numel() - off)
        counts = torch.zeros((max(b, 1),), dtype=torch.int32, device=dev)
        _lib.check(L.se3et_lgr_correspondences(_lib.ptr(score_mat), _lib.i64(score_mat.shape[1]), _lib.ptr(rm), _lib.ptr(sm),
                                               _lib.i64(b), _lib.i64(k), _lib.i64(self.k),
                                               _lib.f32(self.confidence_threshold), wptr, wsize, _lib.ptr(counts),
                                               _lib.stream_ptr()), "lgr_correspondences")
        corr_off = torch.zeros((b + 1,), dtype=torch.int64, device=dev)
        if b:
            corr_off[1:] = torch.cumsum(counts[:b], 0)
        cap = max(1, b * self.k * k)
        out_ref = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        out_src = torch.empty((cap, 3), dtype=torch.float32, device=dev)
        out_sc = torch.empty((cap,), dtype=torch.float32, device=dev)
        out_t = torch.empty((num_pairs, 4, 4), dtype=torch.float32, device=dev)
        _lib.check(L.se3et_lgr_register(_lib.ptr(rp), _lib.ptr(sp), _lib.ptr(patch_offsets), _lib.i64(num_pairs), _lib.i64(b),
                                        _lib.i64(k), _lib.i64(self.k), _lib.f32(self.acceptance_radius),
                                        _lib.i64(self.correspondence_threshold), _lib.i64(self.num_refinement_steps),
                                        wptr, wsize, _lib.ptr(counts), _lib.ptr(corr_off), _lib.ptr(out_ref),
                                        _lib.ptr(out_src), _lib.ptr(out_sc), _lib.ptr(out_t), _lib.stream_ptr()),
                   "lgr_register")
        pair_corr_off = corr_off[patch_offsets]
        total = int(pair_corr_off[-1])  # the one size read-back
        return out_ref[:total], out_src[:total], out_sc[:total], pair_corr_off, out_t

    @torch.no_grad()
    def forward(self, ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat, global_scores=None):
        """Reference signature (one pair): -> ref_corr_points (C, 3), src_corr_points (C, 3), corr_scores (C,),
        estimated_transform (4, 4)."""
        b = score_mat.shape[0]
        off = torch.tensor([0, b], dtype=torch.int64, device=score_mat.device)
        rp, sp, sc, _, t = self.forward_pairs(ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, score_mat, off)
        return rp, sp, sc, t[0]
