"""Host wrappers of the E2PN backbone kernels (include/se3et_b200.h). bf16 activations, fp32 pre-norm values."""
import ctypes

import numpy as np
import torch

from .. import _lib


def builtin_tables():
    """(kidx (15,6), ridx (6,6)) int arrays compiled into the gather kernel."""
    k = (ctypes.c_int32 * 90)()
    r = (ctypes.c_int32 * 36)()
    _lib.check(_lib.lib().se3et_kpconv_tables(k, r), "kpconv_tables")
    return np.array(k, dtype=np.int64).reshape(15, 6), np.array(r, dtype=np.int64).reshape(6, 6)


def kpad_for(cin):
    return (36 * cin + 63) // 64 * 64


def kpconv_gather(q_pts, s_pts, neighbors, x_bf16, kernel_points, kp_extent):
    """-> bf16 (Nq*6, kpad): rows (p, r), columns (class, anchor slot, channel); see se3et_kpconv_gather."""
    _lib.require_cuda(q_pts, s_pts, neighbors, x_bf16, kernel_points)
    assert x_bf16.dtype == torch.bfloat16 and x_bf16.is_contiguous() and x_bf16.dim() == 3 and x_bf16.shape[1] == 6
    assert neighbors.dtype == torch.int64 and neighbors.is_contiguous()
    assert q_pts.dtype == torch.float32 and s_pts.dtype == torch.float32 and q_pts.is_contiguous() and s_pts.is_contiguous()
    assert kernel_points.dtype == torch.float32 and kernel_points.is_contiguous() and kernel_points.shape == (15, 3)
    nq, h = neighbors.shape
    ns, _, cin = x_bf16.shape
    assert s_pts.shape[0] == ns and q_pts.shape[0] == nq
    kpad = kpad_for(cin)
    out = torch.empty((nq * 6, kpad), dtype=torch.bfloat16, device=x_bf16.device)
    _lib.check(_lib.lib().se3et_kpconv_gather(
        _lib.ptr(q_pts), _lib.ptr(s_pts), _lib.ptr(neighbors), _lib.i64(nq), _lib.i64(ns), _lib.i64(h),
        _lib.ptr(x_bf16), _lib.i64(cin), _lib.ptr(kernel_points), _lib.f32(kp_extent), _lib.ptr(out), _lib.i64(kpad),
        _lib.stream_ptr()), "kpconv_gather")
    return out


def groupnorm_stats(y, groups, seg_off, rows_per_point):
    """y: fp32 (rows, C) -> double (nseg, groups, 2) sums / sums of squares per pair."""
    _lib.require_cuda(y, seg_off)
    assert y.dtype == torch.float32 and y.is_contiguous() and y.dim() == 2
    assert seg_off.dtype == torch.int64
    nseg = seg_off.numel() - 1
    stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=y.device)
    _lib.check(_lib.lib().se3et_groupnorm_stats(
        _lib.ptr(y), _lib.i64(y.shape[0]), _lib.i64(y.shape[1]), _lib.i64(groups), _lib.ptr(seg_off), _lib.i64(nseg),
        _lib.i64(rows_per_point), _lib.ptr(stats), _lib.stream_ptr()), "groupnorm_stats")
    return stats


def groupnorm_apply(ya, stats_a, gamma_a, beta_a, groups, seg_off, rows_per_point, slope=0.1, yb=None, stats_b=None,
                    gamma_b=None, beta_b=None, resid=None, out_f32=False, out_bf16=True, eps=1e-5):
    """act(GN(ya) [+ GN(yb)] [+ resid]); slope=1.0 disables the LeakyReLU. Returns (fp32 or None, bf16 or None)."""
    rows, c = ya.shape
    nseg = seg_off.numel() - 1
    of = torch.empty((rows, c), dtype=torch.float32, device=ya.device) if out_f32 else None
    ob = torch.empty((rows, c), dtype=torch.bfloat16, device=ya.device) if out_bf16 else None
    if resid is not None:
        assert resid.dtype == torch.bfloat16 and resid.is_contiguous() and resid.numel() == rows * c
    _lib.check(_lib.lib().se3et_groupnorm_apply(
        _lib.ptr(ya), _lib.ptr(stats_a), _lib.ptr(gamma_a), _lib.ptr(beta_a), _lib.ptr(yb), _lib.ptr(stats_b),
        _lib.ptr(gamma_b), _lib.ptr(beta_b), _lib.ptr(resid), _lib.i64(rows), _lib.i64(c), _lib.i64(groups),
        _lib.ptr(seg_off), _lib.i64(nseg), _lib.i64(rows_per_point), _lib.f32(eps), _lib.f32(slope), _lib.ptr(of),
        _lib.ptr(ob), _lib.stream_ptr()), "groupnorm_apply")
    return of, ob


def maxpool_nbr(x_bf16, neighbors, seg_off=None, seg_width=None):
    """x: bf16 (Ns, A, C) or (Ns, C); neighbors (Nq, H) -> max over neighbours with a zero shadow row.
    seg_off / seg_width (int64 offsets over queries, int32 widths): per-pair effective column count."""
    _lib.require_cuda(x_bf16, neighbors)
    assert x_bf16.dtype == torch.bfloat16 and x_bf16.is_contiguous()
    ns = x_bf16.shape[0]
    width = x_bf16[0].numel() if ns else int(np.prod(x_bf16.shape[1:]))
    nq, h = neighbors.shape
    out = torch.empty((nq,) + tuple(x_bf16.shape[1:]), dtype=torch.bfloat16, device=x_bf16.device)
    nseg = 0
    if seg_width is not None:
        assert seg_width.dtype == torch.int32 and seg_off.dtype == torch.int64
        nseg = seg_width.numel()
        assert seg_off.numel() == nseg + 1
    else:
        seg_off = None
    _lib.check(_lib.lib().se3et_maxpool_nbr(_lib.ptr(x_bf16), _lib.i64(ns), _lib.i64(width), _lib.ptr(neighbors),
                                           _lib.i64(nq), _lib.i64(h), _lib.ptr(seg_off), _lib.ptr(seg_width),
                                           _lib.i64(nseg), _lib.ptr(out), _lib.stream_ptr()), "maxpool_nbr")
    return out


def anchor_max(x_bf16, out=None):
    """(N, A, C) bf16 -> (N, C) max over anchors. `out` may be a column slice of a wider row-major buffer."""
    _lib.require_cuda(x_bf16)
    assert x_bf16.dtype == torch.bfloat16 and x_bf16.is_contiguous() and x_bf16.dim() == 3
    n, a, c = x_bf16.shape
    if out is None:
        out = torch.empty((n, c), dtype=torch.bfloat16, device=x_bf16.device)
    assert out.stride(1) == 1
    _lib.check(_lib.lib().se3et_anchor_max(_lib.ptr(x_bf16), _lib.i64(n), _lib.i64(a), _lib.i64(c), _lib.ptr(out),
                                          _lib.i64(out.stride(0)), _lib.stream_ptr()), "anchor_max")
    return out


def upsample_concat(x_bf16, up_idx, y_bf16):
    """[ xpad[up_idx[:, 0]] | y ] -> (N, C1 + C2) bf16."""
    _lib.require_cuda(x_bf16, up_idx, y_bf16)
    assert x_bf16.is_contiguous() and y_bf16.is_contiguous() and up_idx.dtype == torch.int64
    n = y_bf16.shape[0]
    c1, c2 = x_bf16.shape[1], y_bf16.shape[1]
    out = torch.empty((n, c1 + c2), dtype=torch.bfloat16, device=x_bf16.device)
    _lib.check(_lib.lib().se3et_upsample_concat(
        _lib.ptr(x_bf16), _lib.i64(x_bf16.shape[0]), _lib.i64(c1), _lib.ptr(up_idx), _lib.i64(up_idx.stride(0)),
        _lib.ptr(y_bf16), _lib.i64(c2), _lib.i64(n), _lib.ptr(out), _lib.stream_ptr()), "upsample_concat")
    return out


def kpconv_fused_supported(cin, cout, h):
    """Shapes the fused kernel covers (mirrors se3et_kpconv_fused in csrc/kpconv_fused.cu)."""
    if cin % 16 or cout % 16 or h > 96:
        return False
    # neighbour widths above 48 run as two halves of ceil(h / 2) rounded up to 8 rows per ring stage; 48-row stages
    # leave too little shared memory for the weight ring of a 128-wide tile
    stage_rows = (h + 7) // 8 * 8 if h <= 48 else ((h + 1) // 2 + 7) // 8 * 8
    return stage_rows <= 40 or cout % 128 != 0


def kpconv_fused(q_pts, s_pts, neighbors, x_bf16, w_fused, kernel_points, kp_extent, gn=None, out_bf16=False):
    """KPConvInterSO3.forward in one kernel. x (Ns, 6, Cin) bf16, w_fused (Cout, 36*Cin) bf16 in the fused K order.
    gn = (groups, seg_off) additionally returns the per-pair GroupNorm statistics (double (nseg, groups, 2)).
    -> (fp32 -- bf16 with out_bf16 -- (Nq*6, Cout), stats or None)."""
    _lib.require_cuda(q_pts, s_pts, neighbors, x_bf16, w_fused, kernel_points)
    assert x_bf16.dtype == torch.bfloat16 and x_bf16.is_contiguous() and x_bf16.dim() == 3 and x_bf16.shape[1] == 6
    assert w_fused.dtype == torch.bfloat16 and w_fused.is_contiguous()
    assert neighbors.dtype == torch.int64 and neighbors.is_contiguous()
    assert q_pts.dtype == torch.float32 and s_pts.dtype == torch.float32 and q_pts.is_contiguous() and s_pts.is_contiguous()
    assert kernel_points.dtype == torch.float32 and kernel_points.is_contiguous() and kernel_points.shape == (15, 3)
    nq, h = neighbors.shape
    ns, _, cin = x_bf16.shape
    cout = w_fused.shape[0]
    assert w_fused.shape[1] == 36 * cin and s_pts.shape[0] == ns and q_pts.shape[0] == nq
    out = torch.empty((nq * 6, cout), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x_bf16.device)
    stats, seg_off, groups, nseg = None, None, 0, 0
    if gn is not None:
        groups, seg_off = gn
        assert seg_off.dtype == torch.int64
        nseg = seg_off.numel() - 1
        stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=x_bf16.device)
    _lib.check(_lib.lib().se3et_kpconv_fused(
        _lib.ptr(q_pts), _lib.ptr(s_pts), _lib.ptr(neighbors), _lib.i64(nq), _lib.i64(ns), _lib.i64(h),
        _lib.ptr(x_bf16), _lib.i64(cin), _lib.ptr(w_fused), _lib.i64(cout), _lib.ptr(kernel_points),
        _lib.f32(kp_extent), _lib.ptr(out), ctypes.c_int(1 if out_bf16 else 0), _lib.ptr(stats), _lib.ptr(seg_off),
        _lib.i64(nseg), _lib.i64(groups), _lib.stream_ptr()), "kpconv_fused")
    return out, stats


def kpconv_rows_layout():
    """(src_slot (6, 36), flip (36,)) int arrays: weight slice and channel-half swap per (step anchor, (r, kc))."""
    s = (ctypes.c_int32 * 216)()
    f = (ctypes.c_int32 * 36)()
    _lib.check(_lib.lib().se3et_kpconv_rows_layout(s, f), "kpconv_rows_layout")
    return np.array(s, dtype=np.int64).reshape(6, 36), np.array(f, dtype=np.int64)


def kpconv_rows_supported(cin, cout, h, ns=0):
    """Shapes se3et_kpconv_rows covers (mirrors csrc/kpconv_rows.cu)."""
    return cin % 16 == 0 and cout % 32 == 0 and h <= 48 and ns * 6 * cin // 8 < (1 << 32)


def kpconv_rows(q_pts, s_pts, neighbors, x_bf16, w_rows, kernel_points, kp_extent, out_bf16=False):
    """KPConvInterSO3.forward in one kernel, UMMA rows = points. x (Ns, 6, Cin) bf16, w_rows (Cout, 216*Cin) bf16 in
    step order (KPConvInterSO3._w_rows). -> fp32 (bf16 with out_bf16) (Nq*6, Cout)."""
    _lib.require_cuda(q_pts, s_pts, neighbors, x_bf16, w_rows, kernel_points)
    assert x_bf16.dtype == torch.bfloat16 and x_bf16.is_contiguous() and x_bf16.dim() == 3 and x_bf16.shape[1] == 6
    assert w_rows.dtype == torch.bfloat16 and w_rows.is_contiguous()
    assert neighbors.dtype == torch.int64 and neighbors.is_contiguous()
    assert q_pts.dtype == torch.float32 and s_pts.dtype == torch.float32 and q_pts.is_contiguous() and s_pts.is_contiguous()
    assert kernel_points.dtype == torch.float32 and kernel_points.is_contiguous() and kernel_points.shape == (15, 3)
    nq, h = neighbors.shape
    ns, _, cin = x_bf16.shape
    cout = w_rows.shape[0]
    assert w_rows.shape[1] == 216 * cin and s_pts.shape[0] == ns and q_pts.shape[0] == nq
    out = torch.empty((nq * 6, cout), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=x_bf16.device)
    _lib.check(_lib.lib().se3et_kpconv_rows(
        _lib.ptr(q_pts), _lib.ptr(s_pts), _lib.ptr(neighbors), _lib.i64(nq), _lib.i64(ns), _lib.i64(h),
        _lib.ptr(x_bf16), _lib.i64(cin), _lib.ptr(w_rows), _lib.i64(cout), _lib.ptr(kernel_points),
        _lib.f32(kp_extent), _lib.ptr(out), ctypes.c_int(1 if out_bf16 else 0), _lib.stream_ptr()), "kpconv_rows")
    return out


def groupnorm_double_supported(channels, bf16=False):
    """Row widths se3et_groupnorm_double covers; bf16=True: for a bf16 input (8 columns per thread)."""
    v = channels // 4
    ok = channels % 4 == 0 and 0 < v <= 256 and (v & (v - 1)) == 0
    return ok and (not bf16 or (channels % 8 == 0 and v >= 2))


def groupnorm_double(y, stats1, gamma1, beta1, gamma2, beta2, groups, seg_off, rows_per_point, slope=0.1, eps=1e-5):
    """LeakyReLU(GN_2(LeakyReLU(GN_1(y)))) -> bf16, two streaming passes over y (statistics of the intermediate, then
    apply); see se3et_groupnorm_double."""
    rows, c = y.shape
    nseg = seg_off.numel() - 1
    stats2 = torch.empty((nseg, groups, 2), dtype=torch.float64, device=y.device)
    out = torch.empty((rows, c), dtype=torch.bfloat16, device=y.device)
    L = _lib.lib()
    assert y.is_contiguous() and y.dtype in (torch.float32, torch.bfloat16)
    for apply in (0, 1):
        _lib.check(L.se3et_groupnorm_double(
            _lib.ptr(y), ctypes.c_int(1 if y.dtype == torch.bfloat16 else 0), _lib.ptr(stats1), _lib.ptr(gamma1), _lib.ptr(beta1), _lib.ptr(stats2), _lib.ptr(gamma2),
            _lib.ptr(beta2), _lib.i64(rows), _lib.i64(c), _lib.i64(groups), _lib.ptr(seg_off), _lib.i64(nseg),
            _lib.i64(rows_per_point), _lib.f32(eps), _lib.f32(slope), apply, _lib.ptr(out), _lib.stream_ptr()),
            "groupnorm_double")
    return out


def groupnorm_stats_stream(y, groups, seg_off, rows_per_point):
    """Per-pair GroupNorm statistics (double (nseg, groups, 2)) of fp32 / bf16 y (rows, C) by the streaming kernel
    (se3et_groupnorm_double, apply = 2).  Requires groupnorm_double_supported(C)."""
    rows, c = y.shape
    nseg = seg_off.numel() - 1
    stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=y.device)
    assert y.is_contiguous() and y.dtype in (torch.float32, torch.bfloat16)
    _lib.check(_lib.lib().se3et_groupnorm_double(
        _lib.ptr(y), ctypes.c_int(1 if y.dtype == torch.bfloat16 else 0), _lib.ptr(None), _lib.ptr(None), _lib.ptr(None), _lib.ptr(stats), _lib.ptr(None), _lib.ptr(None),
        _lib.i64(rows), _lib.i64(c), _lib.i64(groups), _lib.ptr(seg_off), _lib.i64(nseg), _lib.i64(rows_per_point),
        _lib.f32(1e-5), _lib.f32(1.0), 2, _lib.ptr(None), _lib.stream_ptr()), "groupnorm_double(stats)")
    return stats


def kpconv_cin1_supported(cin, cout, h):
    return cin == 1 and cout in (32, 64) and h <= 40


def kpconv_cin1(q_pts, s_pts, neighbors, x_bf16, w36, kernel_points, kp_extent, gn=None, lifted=False):
    """First-layer KPConvInterSO3 (one input channel per anchor). x (Ns, 6, 1) bf16 -- or (Ns,) with lifted=True: the
    LiftBlockEPN output, identical for the six anchors -- w36 fp32 (36, Cout).  -> (fp32 (Nq*6, Cout), stats or None)."""
    _lib.require_cuda(q_pts, s_pts, neighbors, x_bf16, w36, kernel_points)
    assert x_bf16.dtype == torch.bfloat16 and x_bf16.is_contiguous()
    assert (x_bf16.dim() == 1) if lifted else (x_bf16.shape[1:] == (6, 1))
    assert w36.dtype == torch.float32 and w36.is_contiguous() and w36.shape[0] == 36
    assert neighbors.dtype == torch.int64 and neighbors.is_contiguous()
    nq, h = neighbors.shape
    ns, cout = x_bf16.shape[0], w36.shape[1]
    out = torch.empty((nq * 6, cout), dtype=torch.float32, device=x_bf16.device)
    stats, seg_off, groups, nseg = None, None, 0, 0
    if gn is not None:
        groups, seg_off = gn
        nseg = seg_off.numel() - 1
        stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=x_bf16.device)
    _lib.check(_lib.lib().se3et_kpconv_cin1(
        _lib.ptr(q_pts), _lib.ptr(s_pts), _lib.ptr(neighbors), _lib.i64(nq), _lib.i64(ns), _lib.i64(h),
        _lib.ptr(x_bf16), _lib.ptr(w36), _lib.i64(cout), _lib.ptr(kernel_points), _lib.f32(kp_extent), _lib.ptr(out),
        _lib.ptr(stats), _lib.ptr(seg_off), _lib.i64(nseg), _lib.i64(groups), ctypes.c_int(1 if lifted else 0),
        _lib.stream_ptr()), "kpconv_cin1")
    return out, stats


def kpconv_lift_supported(cout, h, ns):
    return cout in (32, 64) and h <= 48 and ns < (1 << 31)


def kpconv_lift(q_pts, s_pts, neighbors, f_bf16, w36, kernel_points, kp_extent, gn=None, out_bf16=False):
    """First-layer KPConvInterSO3 on the lifted input (se3et_kpconv_lift): f (Ns,) bf16 = the LiftBlockEPN value of
    every support point, w36 fp32 (36, Cout).  -> (fp32 / bf16 (Nq*6, Cout), stats or None)."""
    _lib.require_cuda(q_pts, s_pts, neighbors, f_bf16, w36, kernel_points)
    assert f_bf16.dtype == torch.bfloat16 and f_bf16.is_contiguous() and f_bf16.dim() == 1
    assert w36.dtype == torch.float32 and w36.is_contiguous() and w36.shape[0] == 36
    assert neighbors.dtype == torch.int64 and neighbors.is_contiguous()
    assert q_pts.dtype == torch.float32 and s_pts.dtype == torch.float32 and q_pts.is_contiguous() and s_pts.is_contiguous()
    nq, h = neighbors.shape
    ns, cout = f_bf16.shape[0], w36.shape[1]
    assert s_pts.shape[0] == ns and q_pts.shape[0] == nq
    out = torch.empty((nq * 6, cout), dtype=torch.bfloat16 if out_bf16 else torch.float32, device=f_bf16.device)
    stats, seg_off, groups, nseg = None, None, 0, 0
    if gn is not None:
        groups, seg_off = gn
        nseg = seg_off.numel() - 1
        stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=f_bf16.device)
    L = _lib.lib()
    ws = _lib.workspace.get(ns * 16 + 512, f_bf16.device)   # >= se3et_kpconv_lift_workspace_bytes(ns)
    _lib.check(L.se3et_kpconv_lift(
        _lib.ptr(q_pts), _lib.ptr(s_pts), _lib.ptr(neighbors), _lib.i64(nq), _lib.i64(ns), _lib.i64(h),
        _lib.ptr(f_bf16), _lib.ptr(w36), _lib.i64(cout), _lib.ptr(kernel_points), _lib.f32(kp_extent), _lib.ptr(out),
        ctypes.c_int(1 if out_bf16 else 0), _lib.ptr(stats), _lib.ptr(seg_off), _lib.i64(nseg), _lib.i64(groups),
        _lib.ptr(ws), _lib.i64(ws.numel()), _lib.stream_ptr()), "kpconv_lift")
    return out, stats
