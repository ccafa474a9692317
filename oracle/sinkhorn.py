"""TEST ORACLE (not a product path): numpy fp32 restatement of LearnableLogOptimalTransport.forward
(geotransformer/modules/sinkhorn/learnable_sinkhorn.py:13-66).  Pinned by tests/test_oracle_sinkhorn.py against
tests/golden/sinkhorn_ref.npz, the output of the reference module itself (tests/golden/make_sinkhorn_golden.py)."""
import numpy as np

INF = np.float32(1e12)


def _lse(x, axis):
    m = x.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(x - m).sum(axis=axis, keepdims=True, dtype=np.float32))).squeeze(axis).astype(np.float32)


def log_optimal_transport(scores, alpha, num_iterations, row_masks=None, col_masks=None):
    """scores (B, M, N) -> (B, M + 1, N + 1), all fp32."""
    scores = np.asarray(scores, np.float32)
    b, m, n = scores.shape
    rm = np.ones((b, m), bool) if row_masks is None else np.asarray(row_masks, bool)
    cm = np.ones((b, n), bool) if col_masks is None else np.asarray(col_masks, bool)
    s = np.full((b, m + 1, n + 1), np.float32(alpha), np.float32)          # dustbin row / column (:40-42)
    s[:, :m, :n] = scores
    pr = np.concatenate([~rm, np.zeros((b, 1), bool)], 1)
    pc = np.concatenate([~cm, np.zeros((b, 1), bool)], 1)
    s[pr[:, :, None] | pc[:, None, :]] = -INF                              # (:43)
    nvr, nvc = rm.sum(1).astype(np.float32), cm.sum(1).astype(np.float32)
    norm = (-np.log(nvr + nvc)).astype(np.float32)                        # (:48)
    log_mu = np.repeat(norm[:, None], m + 1, 1)
    log_mu[:, m] = np.log(nvc) + norm
    log_mu[pr] = -INF
    log_nu = np.repeat(norm[:, None], n + 1, 1)
    log_nu[:, n] = np.log(nvr) + norm
    log_nu[pc] = -INF
    u, v = np.zeros_like(log_mu), np.zeros_like(log_nu)
    for _ in range(num_iterations):                                       # (:13-18)
        u = log_mu - _lse(s + v[:, None, :], 2)
        v = log_nu - _lse(s + u[:, :, None], 1)
    return (s + u[:, :, None] + v[:, None, :] - norm[:, None, None]).astype(np.float32)
