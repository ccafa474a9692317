"""Host wrapper of se3et_gemm_bf16 (tcgen05 GEMM): out = alpha * a @ b.T (+ bias) (ReLU)."""
import ctypes

import torch

from .. import _lib


def _out(spec, m, n, dtype, device):
    """spec: False/None -> no output, True -> allocate, tensor -> write into it (row pitch = stride(0))."""
    if spec is None or spec is False:
        return None
    if spec is True:
        return torch.empty((m, n), dtype=dtype, device=device)
    assert spec.dtype == dtype and spec.shape == (m, n) and spec.stride(1) == 1
    return spec


def linear_bf16(a, w, bias=None, alpha=1.0, relu=False, out_f32=True, out_bf16=False):
    """a: (M, K) bf16 row-major (row pitch a.stride(0)); w: (N, K) bf16 (nn.Linear layout); bias fp32 (N,) or None.
    out_f32 / out_bf16: True to allocate, or a preallocated (M, N) view (columns contiguous) to write into.
    Returns (fp32 or None, bf16 or None) tensors of shape (M, N)."""
    _lib.require_cuda(a, w, bias)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert a.dim() == 2 and w.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.stride(1) == 1 and w.stride(1) == 1
    m, k = a.shape
    n = w.shape[0]
    if n % 16 != 0:
        # the UMMA tile needs N % 16 == 0: zero-pad the weight rows (only reduced-width test configs get here)
        n_pad = (n + 15) // 16 * 16
        w_pad = torch.zeros((n_pad, k), dtype=w.dtype, device=w.device)
        w_pad[:n] = w
        b_pad = None
        if bias is not None:
            b_pad = torch.zeros((n_pad,), dtype=torch.float32, device=w.device)
            b_pad[:n] = bias
        of, ob = linear_bf16(a, w_pad, b_pad, alpha, relu, out_f32, out_bf16)
        return (of[:, :n].contiguous() if of is not None else None, ob[:, :n].contiguous() if ob is not None else None)
    of = _out(out_f32, m, n, torch.float32, a.device)
    ob = _out(out_bf16, m, n, torch.bfloat16, a.device)
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == n
    if m == 0:
        return of, ob
    L = _lib.lib()

    def call(f32, b16):
        ref = f32 if f32 is not None else b16
        _lib.check(L.se3et_gemm_bf16(
            _lib.ptr(a), _lib.i64(a.stride(0) if m > 1 else k), _lib.ptr(w), _lib.i64(w.stride(0) if n > 1 else k),
            _lib.i64(m), _lib.i64(n), _lib.i64(k), _lib.i64(1), _lib.i64(0), _lib.i64(0), _lib.ptr(bias),
            _lib.f32(alpha), int(bool(relu)), _lib.ptr(f32), _lib.ptr(b16), _lib.i64(ref.stride(0) if m > 1 else n),
            _lib.i64(0), _lib.stream_ptr()), "gemm_bf16")

    if of is not None and ob is not None and (m == 1 or of.stride(0) == ob.stride(0)):
        call(of, ob)
    else:
        if of is not None:
            call(of, None)
        if ob is not None:
            call(None, ob)
    return of, ob


def bmm_bf16(a, b, alpha=1.0, out_f32=True, out_bf16=False):
    """Batched: a (B, M, K) bf16, b (B, N, K) bf16 (both contiguous) -> (B, M, N) = alpha * a @ b^T."""
    _lib.require_cuda(a, b)
    assert a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16 and a.is_contiguous() and b.is_contiguous()
    bsz, m, k = a.shape
    n = b.shape[1]
    of = torch.empty((bsz, m, n), dtype=torch.float32, device=a.device) if out_f32 else None
    ob = torch.empty((bsz, m, n), dtype=torch.bfloat16, device=a.device) if out_bf16 else None
    if m == 0 or bsz == 0:
        return of, ob
    L = _lib.lib()
    _lib.check(L.se3et_gemm_bf16(
        _lib.ptr(a), _lib.i64(k), _lib.ptr(b), _lib.i64(k), _lib.i64(m), _lib.i64(n), _lib.i64(k), _lib.i64(bsz),
        _lib.i64(m), _lib.i64(n), _lib.ptr(None), _lib.f32(alpha), 0, _lib.ptr(of), _lib.ptr(ob), _lib.i64(n),
        _lib.i64(m * n), _lib.stream_ptr()), "gemm_bf16")
    return of, ob


def _gn_fusable(n, groups):
    """Mirrors gn_epilogue_ok in csrc/gemm.cu."""
    bn = next((b for b in (256, 128, 64, 32, 16) if n % b == 0), 0)
    if not bn or groups <= 0 or n % groups:
        return False
    cpg = n // groups
    return ((cpg & (cpg - 1)) == 0 and cpg <= 16 or cpg % 16 == 0) and bn // cpg <= 64


def _check_ab(a, w, bias):
    _lib.require_cuda(a, w, bias)
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and a.dim() == 2 and a.shape[1] == w.shape[1]
    assert a.stride(1) == 1 and w.stride(1) == 1
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.is_contiguous() and bias.numel() == w.shape[0]


def linear_gn_stats(a, w, bias, groups, seg_off, rows_per_point, store=True):
    """GroupNorm statistics (double (nseg, groups, 2)) of y = a @ w.T (+ bias), accumulated in the GEMM epilogue;
    store=True also writes y as fp32.  -> (y or None, stats).  Falls back to GEMM + the separate statistics kernel
    for channel/group shapes the epilogue does not cover (still the CUDA path)."""
    from . import e2pn_ops
    m, k = a.shape
    n = w.shape[0]
    if not _gn_fusable(n, groups) or m == 0:
        y, _ = linear_bf16(a, w, bias)
        return y, e2pn_ops.groupnorm_stats(y, groups, seg_off, rows_per_point)
    _check_ab(a, w, bias)
    assert seg_off.dtype == torch.int64
    nseg = seg_off.numel() - 1
    # measured on B200 (scratch/bench_stats.py): widening Linears with a small K are fastest from the Gram matrix of the
    # input (one read of A whatever N is), everything else on the streaming tcgen05 pass (2x the GEMM-epilogue pass for
    # narrowing Linears: 3.4-5.9 TB/s)
    if not store and gram_stats_supported(n, k) and a.stride(0) % 8 == 0:
        return None, linear_gn_stats_gram(a, w, bias, groups, seg_off, rows_per_point)
    if not store and STATS_MODE['stream'] and stream_stats_supported(n, k, groups) and a.stride(0) % 8 == 0:
        return None, linear_gn_stats_stream(a, w, bias, groups, seg_off, rows_per_point)
    y = torch.empty((m, n), dtype=torch.float32, device=a.device) if store else None
    stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=a.device)
    _lib.check(_lib.lib().se3et_gemm_bf16_gnstats(
        _lib.ptr(a), _lib.i64(a.stride(0) if m > 1 else k), _lib.ptr(w), _lib.i64(w.stride(0) if n > 1 else k),
        _lib.i64(m), _lib.i64(n), _lib.i64(k), _lib.ptr(bias), _lib.ptr(y), _lib.i64(n), _lib.ptr(stats),
        _lib.ptr(seg_off), _lib.i64(nseg), _lib.i64(groups), _lib.i64(rows_per_point), _lib.stream_ptr()),
        "gemm_bf16_gnstats")
    return y, stats


def linear_gn_apply(a, w, bias, stats, gamma, beta, eps, slope, groups, seg_off, rows_per_point, resid=None):
    """bf16 LeakyReLU_slope(GroupNorm(a @ w.T + bias) [+ resid]) with the normalisation applied in the GEMM epilogue
    (second pass after linear_gn_stats(..., store=False)).  Requires _gn_fusable(n, groups)."""
    _check_ab(a, w, bias)
    m, k = a.shape
    n = w.shape[0]
    assert _gn_fusable(n, groups) and seg_off.dtype == torch.int64 and stats.dtype == torch.float64
    out = torch.empty((m, n), dtype=torch.bfloat16, device=a.device)
    if m == 0:
        return out
    if resid is not None:
        assert resid.dtype == torch.bfloat16 and resid.is_contiguous() and resid.numel() == m * n
    nseg = seg_off.numel() - 1
    ws = _lib.workspace.get(16 * n * nseg, a.device)
    _lib.check(_lib.lib().se3et_gemm_bf16_gnapply(
        _lib.ptr(a), _lib.i64(a.stride(0) if m > 1 else k), _lib.ptr(w), _lib.i64(w.stride(0) if n > 1 else k),
        _lib.i64(m), _lib.i64(n), _lib.i64(k), _lib.ptr(bias), _lib.ptr(stats), _lib.ptr(gamma), _lib.ptr(beta),
        _lib.f32(eps), _lib.f32(slope), _lib.ptr(resid), _lib.ptr(out), _lib.i64(n), _lib.ptr(seg_off),
        _lib.i64(nseg), _lib.i64(groups), _lib.i64(rows_per_point), _lib.ptr(ws), ctypes.c_size_t(ws.numel()),
        _lib.stream_ptr()), "gemm_bf16_gnapply")
    return out


_DUAL_TILE = {'n': 0}


def linear_gn_apply_dual(a1, w1, bias1, stats1, gamma1, beta1, a2, w2, bias2, stats2, gamma2, beta2, eps, slope, groups,
                         seg_off, rows_per_point):
    """bf16 LeakyReLU_slope(GroupNorm(a1 @ w1.T + bias1) + GroupNorm(a2 @ w2.T + bias2)) in one kernel: the tail of
    ResnetBottleneckBlockEPN (blocks_epn.py:833-852).  Statistics from linear_gn_stats(..., store=False)."""
    _check_ab(a1, w1, bias1)
    _check_ab(a2, w2, bias2)
    m, k1 = a1.shape
    k2 = a2.shape[1]
    n = w1.shape[0]
    assert a2.shape[0] == m and w2.shape[0] == n and n % groups == 0
    assert seg_off.dtype == torch.int64 and stats1.dtype == torch.float64 and stats2.dtype == torch.float64
    out = torch.empty((m, n), dtype=torch.bfloat16, device=a1.device)
    if m == 0:
        return out
    nseg = seg_off.numel() - 1
    ws = _lib.workspace.get(16 * n * nseg, a1.device)
    _lib.check(_lib.lib().se3et_gemm_bf16_gnapply_dual(
        _lib.ptr(a1), _lib.i64(a1.stride(0) if m > 1 else k1), _lib.ptr(w1), _lib.i64(w1.stride(0) if n > 1 else k1),
        _lib.i64(k1), _lib.ptr(bias1), _lib.ptr(stats1), _lib.ptr(gamma1), _lib.ptr(beta1),
        _lib.ptr(a2), _lib.i64(a2.stride(0) if m > 1 else k2), _lib.ptr(w2), _lib.i64(w2.stride(0) if n > 1 else k2),
        _lib.i64(k2), _lib.ptr(bias2), _lib.ptr(stats2), _lib.ptr(gamma2), _lib.ptr(beta2),
        _lib.i64(m), _lib.i64(n), _lib.f32(eps), _lib.f32(slope), _lib.ptr(out), _lib.i64(n), _lib.ptr(seg_off),
        _lib.i64(nseg), _lib.i64(groups), _lib.i64(rows_per_point), ctypes.c_int(_DUAL_TILE['n']),
        _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr()), "gemm_bf16_gnapply_dual")
    return out


def dual_apply_supported(n, k1, k2):
    return n % 32 == 0 and k1 % 8 == 0 and k2 % 8 == 0


_GRAM = {'on': True, 'min_n': 0, 'shared': True}


# A/B switch (tests, measurements): the streaming statistics pass is the product path
STATS_MODE = {'stream': True}


def stream_stats_supported(n, k, groups):
    cpg = n // groups
    return n % 32 == 0 and k % 8 == 0 and n % groups == 0 and (((cpg & (cpg - 1)) == 0 and cpg <= 32) or cpg % 32 == 0)


def linear_gn_stats_stream(a, w, bias, groups, seg_off, rows_per_point):
    """GroupNorm statistics of a @ w.T + bias by the streaming pass (se3et_linear_gnstats_stream): y is never formed."""
    _check_ab(a, w, bias)
    m, k = a.shape
    n = w.shape[0]
    nseg = seg_off.numel() - 1
    stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=a.device)
    _lib.check(_lib.lib().se3et_linear_gnstats_stream(
        _lib.ptr(a), _lib.i64(a.stride(0) if m > 1 else k), _lib.i64(m), _lib.i64(k), _lib.ptr(w),
        _lib.i64(w.stride(0) if n > 1 else k), _lib.i64(n), _lib.ptr(bias), _lib.ptr(seg_off), _lib.i64(nseg),
        _lib.i64(groups), _lib.i64(rows_per_point), _lib.ptr(stats), _lib.stream_ptr()), "linear_gnstats_stream")
    return stats


def gram_stats_supported(n, k):
    """The Gram-matrix statistics pass pays off when the Linear widens (its cost does not depend on n).  Measured on
    B200 (round 2): sending the widening Linears with n <= 256 to the transposed streaming kernel instead (min_n = 256)
    is no faster -- Gram 2.58 -> 0.76 ms but streaming 3.90 -> 5.93 ms per 64 pairs: at K = 32 / 64 a 128-row tile is
    only 8-16 KB and the per-tile hand-offs dominate -- so the Gram pass keeps them."""
    return _GRAM['on'] and k in (32, 64, 128) and n >= 2 * k and n > _GRAM['min_n']


def linear_gn_stats_gram(a, w, bias, groups, seg_off, rows_per_point):
    """GroupNorm statistics of a @ w.T + bias from the Gram matrix of `a` (se3et_linear_gnstats_gram): y is never formed."""
    _check_ab(a, w, bias)
    m, k = a.shape
    n = w.shape[0]
    nseg = seg_off.numel() - 1
    stats = torch.empty((nseg, groups, 2), dtype=torch.float64, device=a.device)
    nbytes = 8 * nseg * (k + 1) * k
    ws = _lib.workspace.get(nbytes, a.device)
    _lib.check(_lib.lib().se3et_linear_gnstats_gram(
        _lib.ptr(a), _lib.i64(a.stride(0) if m > 1 else k), _lib.i64(m), _lib.i64(k), _lib.ptr(w),
        _lib.i64(w.stride(0) if n > 1 else k), _lib.i64(n), _lib.ptr(bias), _lib.ptr(seg_off), _lib.i64(nseg),
        _lib.i64(groups), _lib.i64(rows_per_point), _lib.i64(0), _lib.ptr(ws), ctypes.c_size_t(ws.numel()),
        _lib.ptr(stats), _lib.stream_ptr()), "linear_gnstats_gram")
    return stats


def gram2_supported(k):
    return _GRAM['on'] and _GRAM['shared'] and k in (32, 64, 128)


def linear_gn_stats_gram2(a, w1, bias1, groups1, w2, bias2, groups2, seg_off, rows_per_point):
    """GroupNorm statistics of a @ w1.T + bias1 AND a @ w2.T + bias2 from ONE Gram pass over `a`
    (se3et_linear_gnstats_gram2).  -> (stats1, stats2), double (nseg, groups, 2) each."""
    _check_ab(a, w1, bias1)
    _check_ab(a, w2, bias2)
    m, k = a.shape
    n1, n2 = w1.shape[0], w2.shape[0]
    nseg = seg_off.numel() - 1
    st1 = torch.empty((nseg, groups1, 2), dtype=torch.float64, device=a.device)
    st2 = torch.empty((nseg, groups2, 2), dtype=torch.float64, device=a.device)
    ws = _lib.workspace.get(8 * nseg * (k + 1) * k, a.device)
    _lib.check(_lib.lib().se3et_linear_gnstats_gram2(
        _lib.ptr(a), _lib.i64(a.stride(0) if m > 1 else k), _lib.i64(m), _lib.i64(k),
        _lib.ptr(w1), _lib.i64(w1.stride(0) if n1 > 1 else k), _lib.i64(n1), _lib.ptr(bias1), _lib.i64(groups1), _lib.ptr(st1),
        _lib.ptr(w2), _lib.i64(w2.stride(0) if n2 > 1 else k), _lib.i64(n2), _lib.ptr(bias2), _lib.i64(groups2), _lib.ptr(st2),
        _lib.ptr(seg_off), _lib.i64(nseg), _lib.i64(rows_per_point), _lib.ptr(ws), ctypes.c_size_t(ws.numel()),
        _lib.stream_ptr()), "linear_gnstats_gram2")
    return st1, st2
