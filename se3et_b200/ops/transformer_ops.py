"""Host wrappers of the superpoint-transformer and matching kernels (include/se3et_b200.h)."""
import ctypes

import torch

from .. import _lib


def geo_embed_indices(points, cu, max_cloud, eoff, total_rows, sigma_d, sigma_a, angle_k):
    """points (T,3) fp32, cu (B+1) int64 cloud offsets, eoff (B) int64 row offsets (sum of n_b^2) -> (rows, 4) fp32."""
    _lib.require_cuda(points, cu, eoff)
    assert points.dtype == torch.float32 and points.is_contiguous()
    out = torch.empty((total_rows, 4), dtype=torch.float32, device=points.device)
    _lib.check(_lib.lib().se3et_geo_embed_indices(
        _lib.ptr(points), _lib.ptr(cu), _lib.i64(cu.numel() - 1), _lib.i64(points.shape[0]), _lib.i64(max_cloud),
        _lib.ptr(eoff), _lib.f32(sigma_d), _lib.f32(sigma_a), _lib.i64(angle_k), _lib.ptr(out), _lib.stream_ptr()),
        "geo_embed_indices")
    return out


def geo_embed_project(idx4, w_d, w_a, bias_sum):
    """idx4 (rows,4) fp32; w_d, w_a (C,C) bf16; bias_sum (C,) fp32 = b_d + b_a -> (rows, C) bf16."""
    rows, c = idx4.shape[0], w_d.shape[0]
    out = torch.empty((rows, c), dtype=torch.bfloat16, device=idx4.device)
    _lib.check(_lib.lib().se3et_geo_embed_project(_lib.ptr(idx4), _lib.i64(rows), _lib.i64(c), _lib.ptr(w_d),
                                                 _lib.ptr(w_a), _lib.ptr(bias_sum), _lib.ptr(out), _lib.stream_ptr()),
               "geo_embed_project")
    return out


def geo_embed_lookup(idx4, table_d, table_a, step):
    """idx4 (rows, 4) fp32; table_d (nd, C), table_a (na, C) bf16 tabulated projections with spacing `step`
    -> (rows, C) bf16 = table_d[round(d / step)] + max_k table_a[round(a_k / step)]."""
    rows, c = idx4.shape[0], table_d.shape[1]
    assert table_d.dtype == torch.bfloat16 and table_a.dtype == torch.bfloat16
    assert table_d.is_contiguous() and table_a.is_contiguous() and table_a.shape[1] == c
    out = torch.empty((rows, c), dtype=torch.bfloat16, device=idx4.device)
    _lib.check(_lib.lib().se3et_geo_embed_lookup(
        _lib.ptr(idx4), _lib.i64(rows), _lib.i64(c), _lib.ptr(table_d), _lib.i64(table_d.shape[0]), _lib.ptr(table_a),
        _lib.i64(table_a.shape[0]), _lib.f32(step), _lib.ptr(out), _lib.stream_ptr()), "geo_embed_lookup")
    return out


def gemm_grouped_t(a, a_rows, b, b_rows, groups, max_m, n, n_valid, k, out, alpha=1.0):
    """Grouped GEMM with transposed fp32 stores (see se3et_gemm_grouped_bf16)."""
    _lib.check(_lib.lib().se3et_gemm_grouped_bf16(
        _lib.ptr(a), _lib.i64(a.stride(0)), _lib.i64(a_rows), _lib.ptr(b), _lib.i64(b.stride(0)), _lib.i64(b_rows),
        _lib.ptr(groups), _lib.i64(groups.shape[0]), _lib.i64(max_m), _lib.i64(n), _lib.i64(n_valid), _lib.i64(k),
        _lib.f32(alpha), _lib.ptr(out), _lib.i64(0), 1, _lib.stream_ptr()), "gemm_grouped_bf16")
    return out


def flash_attention(q, q_pt, q_an, k, k_pt, k_an, v, v_pt, v_an, bias, problems, max_q, anchors, heads, head_dim,
                    out, scale=None):
    """q/k/v: bf16 tensors whose data_ptr is the element (point 0, anchor 0, head 0, 0); strides in elements.
    problems: int64 (P,5) device; out: bf16 (rows*anchors, >= heads*head_dim) written at row (q_start+i)*A + a."""
    if scale is None:
        scale = 1.0 / head_dim ** 0.5
    _lib.check(_lib.lib().se3et_flash_attention(
        _lib.ptr(q), _lib.i64(q_pt), _lib.i64(q_an), _lib.ptr(k), _lib.i64(k_pt), _lib.i64(k_an), _lib.ptr(v),
        _lib.i64(v_pt), _lib.i64(v_an), _lib.ptr(bias), _lib.ptr(problems), _lib.i64(problems.shape[0]),
        _lib.i64(max_q), _lib.i64(anchors), _lib.i64(heads), _lib.i64(head_dim), _lib.f32(scale), _lib.ptr(out),
        _lib.i64(out.stride(0)), _lib.stream_ptr()), "flash_attention")
    return out


def add_layernorm(x, resid, resid_div, gamma, beta, eps=1e-5, out_f32=False, out_bf16=True):
    """LayerNorm(x + resid[row // resid_div]); x fp32 (rows, C), resid bf16 or None."""
    rows, c = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    if resid is not None:
        assert resid.dtype == torch.bfloat16 and resid.is_contiguous()
    of = torch.empty((rows, c), dtype=torch.float32, device=x.device) if out_f32 else None
    ob = torch.empty((rows, c), dtype=torch.bfloat16, device=x.device) if out_bf16 else None
    _lib.check(_lib.lib().se3et_add_layernorm(_lib.ptr(x), _lib.ptr(resid), _lib.i64(resid_div), _lib.i64(rows),
                                             _lib.i64(c), _lib.ptr(gamma), _lib.ptr(beta), _lib.f32(eps), _lib.ptr(of),
                                             _lib.ptr(ob), _lib.stream_ptr()), "add_layernorm")
    return of, ob


FUSED_LN = {'on': True}


def linear_add_layernorm_supported(a, w):
    return FUSED_LN['on'] and w.shape[0] == 256 and a.shape[1] % 8 == 0 and a.shape[0] > 0


def linear_add_layernorm(a, w, bias, resid, resid_div, gamma, beta, eps=1e-5):
    """bf16 LayerNorm(resid[row // resid_div] + a @ w.T + bias) in one kernel (se3et_linear_add_layernorm); a bf16
    (rows, K), w bf16 (256, K), resid bf16 (ceil(rows / resid_div), 256)."""
    _lib.require_cuda(a, w, resid, gamma, beta)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and w.dim() == 2
    assert a.stride(1) == 1 and w.stride(1) == 1 and a.shape[1] == w.shape[1]
    assert resid.dtype == torch.bfloat16 and resid.is_contiguous() and resid.shape[1] == w.shape[0]
    assert gamma.dtype == torch.float32 and beta.dtype == torch.float32
    rows, k = a.shape
    n = w.shape[0]
    out = torch.empty((rows, n), dtype=torch.bfloat16, device=a.device)
    _lib.check(_lib.lib().se3et_linear_add_layernorm(
        _lib.ptr(a), _lib.i64(a.stride(0) if rows > 1 else k), _lib.ptr(w), _lib.i64(w.stride(0)), _lib.i64(rows),
        _lib.i64(n), _lib.i64(k), _lib.ptr(bias), _lib.ptr(resid), _lib.i64(resid_div), _lib.ptr(gamma), _lib.ptr(beta),
        _lib.f32(eps), _lib.ptr(out), _lib.stream_ptr()), "linear_add_layernorm")
    return out


def l2_normalize_rows(x, eps=1e-12):
    assert x.dtype == torch.float32 and x.is_contiguous() and x.dim() == 2
    out = torch.empty_like(x)
    _lib.check(_lib.lib().se3et_l2_normalize_rows(_lib.ptr(x), _lib.i64(x.shape[0]), _lib.i64(x.shape[1]),
                                                 _lib.f32(eps), _lib.ptr(out), _lib.stream_ptr()), "l2_normalize_rows")
    return out


def superpoint_matching(ref_feats, src_feats, ref_masks, src_masks, problems, max_ref, max_src, e_total,
                        num_correspondences, dual_normalization=True):
    """Batched SuperPointMatching. ref_feats (Tr, C) / src_feats (Ts, C) fp32 unit rows; masks uint8/bool or None;
    problems int64 (P,5) {ref_start, n_ref, src_start, n_src, e_off}.
    Returns ref_idx (P,k) int64, src_idx (P,k) int64, scores (P,k) fp32, counts (P,) int32 (pair-local indices)."""
    _lib.require_cuda(ref_feats, src_feats, problems)
    assert ref_feats.dtype == torch.float32 and src_feats.dtype == torch.float32
    assert ref_feats.is_contiguous() and src_feats.is_contiguous()
    dev = ref_feats.device
    p = problems.shape[0]
    k = int(num_correspondences)

    def mask8(mk):
        if mk is None:
            return None
        mk = mk.contiguous()
        return mk.view(torch.uint8) if mk.dtype == torch.bool else mk.to(torch.uint8)

    rm, sm = mask8(ref_masks), mask8(src_masks)
    nfloats = ctypes.c_int64(0)
    _lib.check(_lib.lib().se3et_superpoint_matching_workspace_floats(_lib.i64(p), _lib.i64(k), _lib.i64(e_total),
                                                                    ctypes.byref(nfloats)),
               "superpoint_matching_workspace_floats")
    e = torch.empty((max(nfloats.value, 1),), dtype=torch.float32, device=dev)
    rs = torch.empty((max(ref_feats.shape[0], 1),), dtype=torch.float32, device=dev)
    cs = torch.empty((max(src_feats.shape[0], 1),), dtype=torch.float32, device=dev)
    ri = torch.empty((p, k), dtype=torch.int64, device=dev)
    si = torch.empty((p, k), dtype=torch.int64, device=dev)
    sc = torch.empty((p, k), dtype=torch.float32, device=dev)
    cnt = torch.empty((p,), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().se3et_superpoint_matching(
        _lib.ptr(ref_feats), _lib.ptr(src_feats), _lib.i64(ref_feats.shape[1]), _lib.ptr(rm), _lib.ptr(sm),
        _lib.ptr(problems), _lib.i64(p), _lib.i64(max_ref), _lib.i64(max_src), _lib.i64(k),
        int(bool(dual_normalization)), _lib.ptr(e), _lib.i64(e_total), _lib.ptr(rs), _lib.ptr(cs), _lib.ptr(ri),
        _lib.ptr(si), _lib.ptr(sc), _lib.ptr(cnt), _lib.stream_ptr()), "superpoint_matching")
    return ri, si, sc, cnt, e[:max(e_total, 1)]


# ---- SE3ET-E -----------------------------------------------------------------------------------------------------
POSITIVE = {"sq": 0, "softplus": 1, "sigmoid": 2, "relu": 3, "abs": 4}


def anchor_pair_stats(q, q_pt, q_an, k, k_pt, k_an, problems, max_q, anchors, channels, heads, positive):
    """g (P, A, A) fp32 = sum over point pairs of f(head-mean local score); see se3et_anchor_pair_stats."""
    g = torch.empty((problems.shape[0], anchors, anchors), dtype=torch.float32, device=q.device)
    scale = 1.0 / (heads * (channels // heads) ** 0.5)
    _lib.check(_lib.lib().se3et_anchor_pair_stats(
        _lib.ptr(q), _lib.i64(q_pt), _lib.i64(q_an), _lib.ptr(k), _lib.i64(k_pt), _lib.i64(k_an), _lib.ptr(problems),
        _lib.i64(problems.shape[0]), _lib.i64(max_q), _lib.i64(anchors), _lib.i64(channels), _lib.f32(scale),
        int(POSITIVE[positive]), _lib.ptr(g), _lib.stream_ptr()), "anchor_pair_stats")
    return g


def anchor_mix_weights(g, problems, perms, r_soft):
    """-> (w (P, A, A) fp32, attn_r (P, R) fp32 or None)."""
    p, a, _ = g.shape
    w = torch.empty_like(g)
    attn_r = torch.empty((p, perms.shape[0]), dtype=torch.float32, device=g.device) if r_soft else None
    _lib.check(_lib.lib().se3et_anchor_mix_weights(
        _lib.ptr(g), _lib.ptr(problems), _lib.i64(p), _lib.ptr(perms), _lib.i64(perms.shape[0]), _lib.i64(a),
        int(bool(r_soft)), _lib.ptr(w), _lib.ptr(attn_r), _lib.stream_ptr()), "anchor_mix_weights")
    return w, attn_r


def anchor_mix(x, stride_e, stride_n, stride_a, w, cloud_off, anchors, channels, n_points):
    """out (n_points * A, C) bf16 = sum_e w[cloud(n)][a][e] * x[e * stride_e + n * stride_n + a * stride_a + :]."""
    out = torch.empty((n_points * anchors, channels), dtype=torch.bfloat16, device=x.device)
    _lib.check(_lib.lib().se3et_anchor_mix(
        _lib.ptr(x), _lib.i64(stride_e), _lib.i64(stride_n), _lib.i64(stride_a), _lib.ptr(w), _lib.ptr(cloud_off),
        _lib.i64(cloud_off.numel() - 1), _lib.i64(anchors), _lib.i64(channels), _lib.i64(n_points), _lib.ptr(out),
        _lib.stream_ptr()), "anchor_mix")
    return out


def sh_bias_add(points, problems, max_n, u, anchors_mat, heads, c1, bias):
    """bias += l = 1 spherical-harmonics score term; u fp32 (rows, ldu) with 3 values per head."""
    a = anchors_mat.shape[0]
    _lib.check(_lib.lib().se3et_sh_bias_add(
        _lib.ptr(points), _lib.ptr(problems), _lib.i64(problems.shape[0]), _lib.i64(max_n), _lib.ptr(u),
        _lib.i64(u.stride(0)), _lib.ptr(anchors_mat), _lib.i64(a), _lib.i64(heads), _lib.f32(c1), _lib.ptr(bias),
        _lib.stream_ptr()), "sh_bias_add")
    return bias
