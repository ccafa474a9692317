"""TEST INFRASTRUCTURE (never imported by se3et_b200/): numpy restatement of the reference's fine-stage registration.

  weighted_procrustes          geotransformer/modules/registration/procrustes.py:6-73
  LocalGlobalRegistration      geotransformer/modules/geotransformer/local_global_registration.py:49-235
                               (call site experiments/se3eti.3dmatch/model.py:208-224)

Pinned against the unmodified reference modules by tests/golden/make_lgr_golden.py -> tests/golden/lgr_ref.npz
(tests/test_oracle_registration.py).  Where the reference is implementation-defined the oracle fixes an order:
top-k ties -> lower index first; argmax over patch inlier counts -> first maximum."""
import numpy as np


def weighted_procrustes(src, ref, weights=None, weight_thresh=0.0, eps=1e-5):
    """src, ref (N, 3) float; weights (N,) or None -> 4x4 transform mapping src onto ref (procrustes.py:6-73)."""
    src = np.asarray(src, dtype=np.float64).reshape(-1, 3)
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, 3)
    w = np.ones(len(src)) if weights is None else np.asarray(weights, dtype=np.float64).copy()
    w[w < weight_thresh] = 0.0
    w = w / (w.sum() + eps)
    cs = (src * w[:, None]).sum(0)
    cr = (ref * w[:, None]).sum(0)
    H = (src - cs).T @ (w[:, None] * (ref - cr))
    U, _, Vt = np.linalg.svd(H)
    V = Vt.T
    d = np.sign(np.linalg.det(V @ U.T))
    R = V @ np.diag([1.0, 1.0, d]) @ U.T
    t = cr - R @ cs
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def _topk_mask(score, k, axis):
    """Boolean mask of the k largest entries along `axis`, ties broken towards the lower index."""
    order = np.argsort(-score, axis=axis, kind="stable")
    idx = np.take(order, np.arange(min(k, score.shape[axis])), axis=axis)
    mask = np.zeros(score.shape, dtype=bool)
    np.put_along_axis(mask, idx, True, axis=axis)
    return mask


def correspondence_matrix(score_mat, ref_masks, src_masks, k, confidence_threshold, mutual=True):
    """score_mat (B, K, K) = exp(log scores) (local_global_registration.py:49-84)."""
    ref_corr = _topk_mask(score_mat, k, 2) & (score_mat > confidence_threshold)
    src_corr = _topk_mask(score_mat, k, 1) & (score_mat > confidence_threshold)
    corr = (ref_corr & src_corr) if mutual else (ref_corr | src_corr)
    return corr & (ref_masks[:, :, None] & src_masks[:, None, :])


def local_global_registration(ref_knn_points, src_knn_points, ref_knn_masks, src_knn_masks, log_scores, k=3,
                              acceptance_radius=0.1, mutual=True, confidence_threshold=0.05,
                              correspondence_threshold=3, num_refinement_steps=5):
    """-> (ref_corr_points (C, 3), src_corr_points (C, 3), corr_scores (C,), transform (4, 4)); log_scores is the
    (B, K, K) block of the optimal-transport output without the dustbin row / column (model.py:208-224)."""
    score = np.exp(np.asarray(log_scores, dtype=np.float32))
    corr = correspondence_matrix(score, np.asarray(ref_knn_masks, bool), np.asarray(src_knn_masks, bool), k,
                                 confidence_threshold, mutual)
    score = score * corr.astype(np.float32)
    b, r, s = np.nonzero(corr)
    ref_pts = np.asarray(ref_knn_points, dtype=np.float32)[b, r]
    src_pts = np.asarray(src_knn_points, dtype=np.float32)[b, s]
    sc = score[b, r, s]

    def apply(T, p):
        return p.astype(np.float64) @ T[:3, :3].T + T[:3, 3]

    def rescored(T):
        res = np.linalg.norm(ref_pts - apply(T, src_pts), axis=1)
        return sc * (res < acceptance_radius)

    chunks = []
    if len(b):
        cut = np.flatnonzero(b[1:] != b[:-1]) + 1
        bounds = np.concatenate([[0], cut, [len(b)]])
        chunks = [(x, y) for x, y in zip(bounds[:-1], bounds[1:]) if y - x >= correspondence_threshold]
    if chunks:
        best, best_count = None, -1
        for x, y in chunks:
            T = weighted_procrustes(src_pts[x:y], ref_pts[x:y], sc[x:y])
            res = np.linalg.norm(ref_pts - apply(T, src_pts), axis=1)
            inl = res < acceptance_radius
            if inl.sum() > best_count:
                best, best_count = inl, int(inl.sum())
        cur = sc * best
    else:
        cur = rescored(weighted_procrustes(src_pts, ref_pts, sc))
    T = weighted_procrustes(src_pts, ref_pts, cur)
    for _ in range(num_refinement_steps - 1):
        T = weighted_procrustes(src_pts, ref_pts, rescored(T))
    return ref_pts, src_pts, sc, T.astype(np.float32)
