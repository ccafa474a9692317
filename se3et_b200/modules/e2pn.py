"""E2PN equivariant KPConv blocks on the CUDA path -- drop-in mirrors of
geotransformer/modules/e2pn/blocks_epn.py (same class names, constructor arguments, forward signatures and
state_dict keys: `weights`, `kernel_points`, `anchors`, `quotient_anchors`, `kidx_rot`, `ridx_rot`, `mlp.*`,
`norm.norm.*`), plus the invariant decoder pieces of geotransformer/modules/kpconv/modules.py.

Execution model: activations travel between blocks as bf16 (N, A, C) tensors; every Linear / KPConv contraction is
one tcgen05 GEMM producing fp32 pre-norm values; GroupNorm statistics are per point-cloud PAIR (`seg`: int64
offsets into the stacked point axis; None = the whole tensor is one pair, exactly the reference's behaviour).
fp32 inputs are accepted everywhere and rounded to bf16 once.
"""
import math
import threading

import numpy as np
import torch
import torch.nn as nn
from torch.nn.parameter import Parameter

from .. import _lib
from ..ops import e2pn_ops as K
from ..ops.gemm import (_gn_fusable, dual_apply_supported, gram2_supported, linear_bf16, linear_gn_apply,
                        linear_gn_apply_dual, linear_gn_stats, linear_gn_stats_gram2)
from . import octahedral


# switches for A/B measurements and tests (the defaults are the product path)
# 'cin1_kernel': the CUDA-core first-layer kernel (csrc/kpconv.cu) measures slower than gather + GEMM on B200
# (9.6 vs 8.4 ms per 64 pairs), so it is off by default and only exercised by the tests
_GFLAGS = {'fused_kpconv': True, 'two_pass_unary': True, 'double_norm': True, 'cin1_kernel': False, 'dual_apply': True, 'lifted_kernel': True, 'conv_stats_stream': True,
           'rows_kpconv': True, 'rows_max_cout': 128,
           # pre-norm conv outputs in bf16 between the conv kernel and the double GroupNorm (inference path only)
           'conv_bf16': True,
           # the lifted first layer on se3et_kpconv_lift (thread per point + mma.sync) instead of se3et_kpconv_cin1
           'lift_kernel': True}


def _gn_fusable_fused(cout, groups):
    """GroupNorm shapes the fused KPConv epilogue covers (same rule as the GEMM epilogue, tile width <= 128)."""
    if cout % groups:
        return False
    bn = next(b for b in (128, 64, 32, 16) if cout % b == 0)
    cpg, chunk = cout // groups, min(bn, 32)
    return ((cpg & (cpg - 1)) == 0 and cpg <= chunk or cpg % chunk == 0) and bn // cpg <= 64


def _act(x):
    return x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)


def _seg(seg, n_points, device):
    if seg is None:
        return torch.tensor([0, n_points], dtype=torch.int64, device=device)
    return seg


class _Bf16Cache:
    """bf16 copy of an fp32 parameter, refreshed when the parameter changes (in-place updates bump _version).

    Launch sequences run on several host threads, one CUDA stream each (GeoTransformer.forward_stacked_concurrent): the
    conversion is enqueued once, under a lock, on the stream of whichever thread gets there first; every other stream
    waits on the conversion's event the first time it reads the copy."""

    def __init__(self):
        self._key, self._val, self._event, self._synced = None, None, None, set()
        self._lock = threading.Lock()

    def get(self, p, transform=None, extra=None, dtype=torch.bfloat16):
        """`extra`: a second parameter the transform reads (its version joins the key); `dtype`: of the cached copy."""
        key = (p.data_ptr(), p._version, p.device)
        if extra is not None:
            key = key + (extra.data_ptr(), extra._version)
        if key != self._key:
            with self._lock:
                if key != self._key:
                    with torch.no_grad():
                        v = p.detach() if transform is None else transform(p.detach())
                        val = v.to(dtype).contiguous()
                    ev, synced = None, set()
                    if val.is_cuda:
                        st = torch.cuda.current_stream(val.device)
                        ev = torch.cuda.Event()
                        ev.record(st)
                        synced.add(st.cuda_stream)
                    self._val, self._event, self._synced = val, ev, synced
                    self._key = key  # last: whoever sees the key sees the value and its event
        val = self._val
        if val.is_cuda and _lib._raw_stream(val.device.index) not in self._synced:   # raw handle: this is a hot path
            st = torch.cuda.current_stream(val.device)
            with self._lock:
                if self._event is not None:
                    st.wait_event(self._event)
                self._synced.add(st.cuda_stream)
        return val


class GroupNormEPN(nn.Module):
    """blocks_epn.py:684-701. forward(x): x (N, A, C) -> normalised (N, A, C)."""

    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups = num_groups
        self.num_channels = num_channels
        self.norm = nn.GroupNorm(self.num_groups, self.num_channels)

    def fused(self, y, seg, rows_per_point, slope, out_f32=False, out_bf16=True, other=None, resid=None, stats=None):
        """y: fp32 (rows, C) pre-norm (stats: its statistics when the producing GEMM already accumulated them).
        other = (y_b, GroupNormEPN_b, stats_b or None) adds a second normalised operand."""
        if stats is None:
            stats = K.groupnorm_stats(y, self.num_groups, seg, rows_per_point)
        kw = {}
        if other is not None:
            yb, nb, sb = other
            if sb is None:
                sb = K.groupnorm_stats(yb, nb.num_groups, seg, rows_per_point)
            kw = dict(yb=yb, stats_b=sb, gamma_b=nb.norm.weight, beta_b=nb.norm.bias)
        return K.groupnorm_apply(y, stats, self.norm.weight, self.norm.bias, self.num_groups, seg, rows_per_point,
                                 slope=slope, resid=resid, out_f32=out_f32, out_bf16=out_bf16, eps=self.norm.eps, **kw)

    def forward(self, x, seg=None):
        shape = x.shape
        rpp = shape[1] if x.dim() == 3 else 1
        y = x.reshape(-1, shape[-1]).float().contiguous()
        f32 = x.dtype == torch.float32
        of, ob = self.fused(y, _seg(seg, shape[0], x.device), rpp, slope=1.0, out_f32=f32, out_bf16=not f32)
        return (of if f32 else ob).reshape(shape).squeeze()


class KPConvInterSO3(nn.Module):
    """blocks_epn.py:18-552, for the configuration SE3ET uses (kanchor 6, quotient_factor 4, 15 kernel points,
    'linear' influence, 'sum' aggregation, non_sep_conv, rot_by_permute, fixed 'center')."""

    def __init__(self, kernel_size, kanchor, in_channels, out_channels, KP_extent, radius, KP_influence='linear',
                 aggregation_mode='sum', deformable=False, modulated=False, epn_kernel=False, equiv_mode_kp=False,
                 non_sep_conv=False, rot_by_permute=False, fixed_kernel_points='center', quotient_factor=1,
                 ignore_steer_constraint=False, gather_by_idxing=False):
        super().__init__()
        ok = (kernel_size == 15 and kanchor == 6 and quotient_factor == 4 and KP_influence == 'linear' and
              aggregation_mode == 'sum' and non_sep_conv and rot_by_permute and fixed_kernel_points == 'center' and
              not deformable and not epn_kernel and not ignore_steer_constraint)
        if not ok:
            raise NotImplementedError("se3et_b200.KPConvInterSO3 supports the SE3ET configuration only "
                                      "(K=15, kanchor=6, quotient_factor=4, linear/sum, non_sep_conv, rot_by_permute)")
        self.kanchor, self.K, self.K_real = kanchor, kernel_size, octahedral.K_REAL
        self.in_channels, self.out_channels = in_channels, out_channels
        self.radius, self.KP_extent = radius, KP_extent
        self.KP_influence, self.aggregation_mode = KP_influence, aggregation_mode
        self.quotient_factor = quotient_factor
        t = octahedral.tables()
        self.kernel_points = Parameter(torch.tensor(t["kp_unit"] * 0.7 * radius, dtype=torch.float32),
                                       requires_grad=False)
        self.quotient_anchors = Parameter(torch.tensor(t["quotient"], dtype=torch.float32), requires_grad=False)
        self.anchors = Parameter(torch.tensor(t["anchors"], dtype=torch.float32), requires_grad=False)
        kidx = torch.tensor(t["kidx"], dtype=torch.int64)  # (K, R)
        ridx = torch.tensor(t["ridx"], dtype=torch.int64)  # (A, R)
        self.register_buffer('kidx_rot', kidx[:, None, :].expand(-1, kanchor, -1).contiguous())
        self.register_buffer('ridx_rot', ridx[None, :, :].expand(kernel_size, -1, -1).contiguous())
        self.weights = Parameter(torch.zeros((self.K_real, kanchor, in_channels, out_channels), dtype=torch.float32),
                                 requires_grad=True)
        self.reset_parameters()
        self._w_cache = _Bf16Cache()
        self._wf_cache = _Bf16Cache()
        self._wr_cache = _Bf16Cache()
        self._tables_checked = False

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weights, a=math.sqrt(5))

    def _check_tables(self):
        if not self._tables_checked:
            kidx, ridx = K.builtin_tables()
            if not (np.array_equal(self.kidx_rot[:, 0, :].cpu().numpy(), kidx) and
                    np.array_equal(self.ridx_rot[0].cpu().numpy(), ridx)):
                raise RuntimeError("KPConvInterSO3: kidx_rot/ridx_rot differ from the tables compiled into the kernel")
            self._tables_checked = True

    def _w_flat(self):
        """(Cout, kpad) bf16, K-major: W[kc, a', c, d] -> row d, column (kc*6 + a')*Cin + c (zero padded)."""
        kpad = K.kpad_for(self.in_channels)

        def tr(w):
            flat = w.reshape(-1, self.out_channels).t()
            if kpad != flat.shape[1]:
                flat = torch.nn.functional.pad(flat, (0, kpad - flat.shape[1]))
            return flat
        return self._w_cache.get(self.weights, tr)

    def _w_fused(self):
        """(Cout, 36*Cin) bf16 for the fused kernel: K index = (chunk*36 + kc*6 + a')*16 + c, input channel chunk*16+c."""
        def tr(w):
            cin, cout = self.in_channels, self.out_channels
            return w.reshape(36, cin // 16, 16, cout).permute(3, 1, 0, 2).reshape(cout, 36 * cin)
        return self._wf_cache.get(self.weights, tr)

    def _w_rows(self):
        """(Cout, 216*Cin) bf16 for se3et_kpconv_rows: K index = ((chunk*6 + a)*36 + r*6 + kc)*16 + c', value
        W[kc, ridx[a][r], chunk*16 + (c' ^ 8*flip[r*6 + kc]), d] (tables from the library: se3et_kpconv_rows_layout)."""
        def tr(w):
            cin, cout = self.in_channels, self.out_channels
            slot, flip = K.kpconv_rows_layout()
            slot = torch.as_tensor(slot, device=w.device)                      # (6 a, 36 t)
            c = torch.arange(16, device=w.device)
            cidx = c[None, :] ^ (8 * torch.as_tensor(flip, device=w.device))[:, None]   # (36 t, 16)
            w4 = w.reshape(36, cin // 16, 16, cout)[slot]                      # (6, 36, nch, 16, cout)
            w4 = torch.gather(w4, 3, cidx[None, :, None, :, None].expand(6, 36, cin // 16, 16, cout))
            return w4.permute(4, 2, 0, 1, 3).reshape(cout, 216 * cin)
        return self._wr_cache.get(self.weights, tr)

    def _rows_ok(self, neighb_inds, ns):
        return _GFLAGS['rows_kpconv'] and neighb_inds.shape[0] > 0 and self.out_channels <= _GFLAGS['rows_max_cout'] \
            and K.kpconv_rows_supported(self.in_channels, self.out_channels, neighb_inds.shape[1], ns)

    def _conv(self, q_pts, s_pts, neighb_inds, x, out_bf16=False):
        """fp32 (bf16 with out_bf16) (Nq*6, Cout) by the fused kernels (caller checked _fused_ok)."""
        if self._rows_ok(neighb_inds, s_pts.shape[0]):
            return K.kpconv_rows(q_pts, s_pts, neighb_inds.contiguous(), _act(x).contiguous(), self._w_rows(),
                                 self.kernel_points, self.KP_extent, out_bf16=out_bf16)
        y, _ = K.kpconv_fused(q_pts, s_pts, neighb_inds.contiguous(), _act(x).contiguous(), self._w_fused(),
                              self.kernel_points, self.KP_extent, out_bf16=out_bf16)
        return y

    def _fused_ok(self, neighb_inds):
        """One of the one-kernel paths (se3et_kpconv_rows / se3et_kpconv_fused) covers this shape."""
        return _GFLAGS['fused_kpconv'] and neighb_inds.shape[0] > 0 and (
            self._rows_ok(neighb_inds, 0) or
            K.kpconv_fused_supported(self.in_channels, self.out_channels, neighb_inds.shape[1]))

    def forward(self, q_pts, s_pts, neighb_inds, x):
        """-> fp32 (Nq, A, Cout), pre-norm (blocks_epn.py:454-546)."""
        self._check_tables()
        if self._fused_ok(neighb_inds):
            return self._conv(q_pts, s_pts, neighb_inds, x).view(-1, self.kanchor, self.out_channels)
        a = K.kpconv_gather(q_pts, s_pts, neighb_inds.contiguous(), _act(x).contiguous(), self.kernel_points,
                            self.KP_extent)
        y, _ = linear_bf16(a, self._w_flat())
        return y.view(-1, self.kanchor, self.out_channels)

    def forward_stats(self, q_pts, s_pts, neighb_inds, x, groups, seg, allow_bf16=False):
        """forward() plus the per-pair GroupNorm statistics of its output (accumulated in the GEMM epilogue).
        allow_bf16: the caller (the two back-to-back norms, _conv_double_norm) also reads a bf16 pre-norm tensor; the
        one-kernel convolutions then write bf16 and the three streaming passes that follow move half the bytes."""
        self._check_tables()
        cin1_ok = neighb_inds.shape[0] > 0 and s_pts.shape[0] > 0 and K.kpconv_cin1_supported(
            self.in_channels, self.out_channels, neighb_inds.shape[1])
        if cin1_ok and _GFLAGS['lifted_kernel'] and x.dim() == 3 and x.stride(1) == 0:
            # LiftBlockEPN output (an expand over the anchor axis): anchor-constant input, 16 products per point
            w36 = self.weights.detach().reshape(36, self.out_channels).float().contiguous()
            if _GFLAGS['lift_kernel'] and K.kpconv_lift_supported(self.out_channels, neighb_inds.shape[1], s_pts.shape[0]):
                # round 2: a thread per point + mma.sync for the 16 -> 6 * Cout product (csrc/kpconv_lift.cu)
                return K.kpconv_lift(q_pts, s_pts, neighb_inds.contiguous(), _act(x[:, 0, 0]).contiguous(), w36,
                                     self.kernel_points, self.KP_extent, gn=(groups, seg),
                                     out_bf16=allow_bf16 and _GFLAGS['conv_bf16'] and
                                     K.groupnorm_double_supported(self.out_channels, bf16=True))
            return K.kpconv_cin1(q_pts, s_pts, neighb_inds.contiguous(), _act(x[:, 0, 0]).contiguous(), w36,
                                 self.kernel_points, self.KP_extent, gn=(groups, seg), lifted=True)
        if _GFLAGS['cin1_kernel'] and cin1_ok:
            w36 = self.weights.detach().reshape(36, self.out_channels).float().contiguous()
            return K.kpconv_cin1(q_pts, s_pts, neighb_inds.contiguous(), _act(x).contiguous(), w36, self.kernel_points,
                                 self.KP_extent, gn=(groups, seg))
        if self._fused_ok(neighb_inds) and _gn_fusable_fused(self.out_channels, groups):
            if _GFLAGS['conv_stats_stream'] and K.groupnorm_double_supported(self.out_channels):
                # the statistics as a streaming pass over the (small) conv output: in the fused kernel's epilogue they
                # sit on the producers' critical path (4-10 % of that kernel), here they cost one read of y
                y = self._conv(q_pts, s_pts, neighb_inds, x, out_bf16=allow_bf16 and _GFLAGS['conv_bf16'] and
                               K.groupnorm_double_supported(self.out_channels, bf16=True))
                return y, K.groupnorm_stats_stream(y, groups, seg, self.kanchor)
            return K.kpconv_fused(q_pts, s_pts, neighb_inds.contiguous(), _act(x).contiguous(), self._w_fused(),
                                  self.kernel_points, self.KP_extent, gn=(groups, seg))
        a = K.kpconv_gather(q_pts, s_pts, neighb_inds.contiguous(), _act(x).contiguous(), self.kernel_points,
                            self.KP_extent)
        return linear_gn_stats(a, self._w_flat(), None, groups, seg, self.kanchor)

    def __repr__(self):
        return 'KPConvInterSO3(radius: {:.2f}, extent: {:.2f}, in_feat: {:d}, out_feat: {:d})'.format(
            self.radius, self.KP_extent, self.in_channels, self.out_channels)


class UnaryBlockEPN(nn.Module):
    """blocks_epn.py:639-665: Linear -> GroupNormEPN -> LeakyReLU(0.1) (unless no_relu)."""

    def __init__(self, in_dim, out_dim, group_norm, bn_momentum, no_relu=False):
        super().__init__()
        self.bn_momentum, self.no_relu = bn_momentum, no_relu
        self.in_dim, self.out_dim = in_dim, out_dim
        self.mlp = nn.Linear(in_dim, out_dim)
        self.norm = GroupNormEPN(group_norm, out_dim)
        self.leaky_relu = nn.LeakyReLU(0.1)
        self._w_cache = _Bf16Cache()

    def pre_norm(self, x, seg, rows_per_point, store=True):
        """fp32 (N*A, Cout) Linear output (None when store=False) and its per-pair GroupNorm statistics."""
        x2 = _act(x).reshape(-1, self.in_dim)
        return linear_gn_stats(x2, self._w_cache.get(self.mlp.weight), self.mlp.bias, self.norm.num_groups, seg,
                               rows_per_point, store=store)

    def two_pass_ok(self):
        return _GFLAGS['two_pass_unary'] and _gn_fusable(self.out_dim, self.norm.num_groups)

    def apply(self, x, stats, seg, rows_per_point, slope, resid=None):
        """Second pass: the Linear recomputed with GroupNorm (+ resid) (+ LeakyReLU) in the GEMM epilogue -> bf16."""
        x2 = _act(x).reshape(-1, self.in_dim)
        return linear_gn_apply(x2, self._w_cache.get(self.mlp.weight), self.mlp.bias, stats, self.norm.norm.weight,
                               self.norm.norm.bias, self.norm.norm.eps, slope, self.norm.num_groups, seg,
                               rows_per_point, resid=resid)

    def forward(self, x, batch=None, seg=None):
        n, a = x.shape[0], x.shape[1]
        seg = _seg(seg, n, x.device)
        slope = 1.0 if self.no_relu else 0.1
        if self.two_pass_ok() and n > 0:
            x = _act(x).contiguous()
            _, stats = self.pre_norm(x, seg, a, store=False)
            return self.apply(x, stats, seg, a, slope).view(n, a, self.out_dim)
        y, stats = self.pre_norm(x, seg, a)
        _, out = self.norm.fused(y, seg, a, slope=slope, stats=stats)
        return out.view(n, a, self.out_dim)


class LastUnaryBlockEPN(nn.Module):
    """blocks_epn.py:668-681."""

    def __init__(self, in_dim, out_dim, bias=True):
        super().__init__()
        self.in_dim, self.out_dim = in_dim, out_dim
        self.mlp = nn.Linear(in_dim, out_dim, bias=bias)
        self._w_cache = _Bf16Cache()

    def forward(self, x):
        y, _ = linear_bf16(_act(x).reshape(-1, self.in_dim), self._w_cache.get(self.mlp.weight), self.mlp.bias)
        return y.view(*x.shape[:-1], self.out_dim)


class KPConvInterSO3Block(nn.Module):
    """blocks_epn.py:703-743: conv -> GroupNormEPN -> LeakyReLU."""

    def __init__(self, block_name, in_dim, out_dim, radius, sigma, group_norm, config):
        super().__init__()
        self.block_name, self.in_dim, self.out_dim = block_name, in_dim, out_dim
        self.conv = KPConvInterSO3(config.num_kernel_points, config.kanchor, in_dim, out_dim, sigma, radius,
                                   config.KP_influence, config.aggregation_mode, epn_kernel=config.epn_kernel,
                                   equiv_mode_kp=config.equiv_mode_kp, non_sep_conv=config.non_sep_conv,
                                   rot_by_permute=config.rot_by_permute,
                                   fixed_kernel_points=config.fixed_kernel_points,
                                   quotient_factor=config.quotient_factor,
                                   ignore_steer_constraint=config.ignore_steer_constraint,
                                   gather_by_idxing=config.gather_by_idxing)
        self.norm = GroupNormEPN(group_norm, out_dim)
        self.leaky_relu = nn.LeakyReLU(0.1)

    def fused(self, x, q_pts, s_pts, neighb_inds, seg, out_f32):
        y, stats = self.conv.forward_stats(q_pts, s_pts, neighb_inds, x, self.norm.num_groups, seg)
        return self.norm.fused(y, seg, self.conv.kanchor, slope=0.1, out_f32=out_f32, out_bf16=not out_f32,
                               stats=stats)

    def forward(self, x, q_pts, s_pts, neighb_inds, seg=None):
        _, out = self.fused(x, q_pts, s_pts, neighb_inds, _seg(seg, q_pts.shape[0], q_pts.device), False)
        return out.view(-1, self.conv.kanchor, self.out_dim)


def _conv_double_norm(interso3, norm2, x, q_pts, s_pts, neighb_inds, seg):
    """conv -> GroupNormEPN -> LeakyReLU (KPConvInterSO3Block) -> GroupNormEPN -> LeakyReLU (enclosing block), bf16 out.
    The conv kernel delivers the first statistics; the intermediate activation is never written."""
    n1 = interso3.norm
    c = interso3.out_dim
    if _GFLAGS['double_norm'] and K.groupnorm_double_supported(c) and n1.num_groups == norm2.num_groups and \
            n1.norm.eps == norm2.norm.eps and q_pts.shape[0] > 0:
        y, stats = interso3.conv.forward_stats(q_pts, s_pts, neighb_inds, x, n1.num_groups, seg, allow_bf16=True)
        return K.groupnorm_double(y.view(-1, c), stats, n1.norm.weight, n1.norm.bias, norm2.norm.weight,
                                  norm2.norm.bias, n1.num_groups, seg, 6, slope=0.1, eps=n1.norm.eps)
    f, _ = interso3.fused(x, q_pts, s_pts, neighb_inds, seg, out_f32=True)
    _, out = norm2.fused(f, seg, 6, slope=0.1)
    return out


class SimpleBlockEPN(nn.Module):
    """blocks_epn.py:770-796: interso3 block, then a second GroupNormEPN + LeakyReLU."""

    def __init__(self, block_name, in_dim, out_dim, radius, sigma, group_norm, config):
        super().__init__()
        if not config.non_sep_conv:
            raise NotImplementedError("separable (intra-SO3) convolution is not part of the SE3ET path")
        self.block_name, self.in_dim, self.out_dim = block_name, in_dim, out_dim
        self.non_sep_conv = config.non_sep_conv
        self.interso3 = KPConvInterSO3Block(block_name, in_dim, out_dim, radius, sigma, group_norm, config)
        self.norm = GroupNormEPN(group_norm, out_dim)
        self.leaky_relu = nn.LeakyReLU(0.1)

    def forward(self, x, q_pts, s_pts, neighb_inds, seg=None):
        seg = _seg(seg, q_pts.shape[0], q_pts.device)
        out = _conv_double_norm(self.interso3, self.norm, x, q_pts, s_pts, neighb_inds, seg)
        return out.view(-1, 6, self.out_dim)


class ResnetBottleneckBlockEPN(nn.Module):
    """blocks_epn.py:798-852."""

    def __init__(self, block_name, in_dim, out_dim, radius, sigma, group_norm, config):
        super().__init__()
        if not config.non_sep_conv:
            raise NotImplementedError("separable (intra-SO3) convolution is not part of the SE3ET path")
        self.bn_momentum = config.batch_norm_momentum
        self.block_name, self.in_dim, self.out_dim = block_name, in_dim, out_dim
        self.relu_end = True
        self.leaky_relu = nn.LeakyReLU(0.1)
        self.non_sep_conv = config.non_sep_conv
        if in_dim != out_dim // 4:
            self.unary1 = UnaryBlockEPN(in_dim, out_dim // 4, group_norm, self.bn_momentum)
        else:
            self.unary1 = nn.Identity()
        self.interso3 = KPConvInterSO3Block(block_name, out_dim // 4, out_dim // 4, radius, sigma, group_norm, config)
        self.norm = GroupNormEPN(group_norm, out_dim // 4)
        self.unary2 = UnaryBlockEPN(out_dim // 4, out_dim, group_norm, self.bn_momentum, no_relu=self.relu_end)
        if in_dim != out_dim:
            self.skip_conv = UnaryBlockEPN(in_dim, out_dim, group_norm, self.bn_momentum, no_relu=self.relu_end)
        else:
            self.skip_conv = nn.Identity()

    def forward(self, x, q_pts, s_pts, neighb_inds, seg=None, s_seg=None, sub_width=None):
        """seg: pair offsets of the query level; s_seg: of the support level (needed only when strided);
        sub_width: per-pair column count of the pooling matrix when several pairs share neighb_inds."""
        nq = q_pts.shape[0]
        seg = _seg(seg, nq, q_pts.device)
        x = _act(x).contiguous()
        skip = x
        st_skip = None   # statistics of the shortcut Linear when they come from the shared Gram pass below
        has_skip_conv = isinstance(self.skip_conv, UnaryBlockEPN)
        if isinstance(self.unary1, UnaryBlockEPN):
            s_seg = _seg(s_seg, s_pts.shape[0], s_pts.device) if 'strided' in self.block_name else seg
            u1, sc = self.unary1, self.skip_conv
            if ('strided' not in self.block_name and has_skip_conv and nq > 0 and gram2_supported(self.in_dim)
                    and u1.two_pass_ok() and sc.two_pass_ok() and x.stride(-1) == 1):
                # the block input feeds unary1 AND the shortcut Linear: one Gram pass over x gives both statistics
                x2 = x.reshape(-1, self.in_dim)
                st1, st_skip = linear_gn_stats_gram2(
                    x2, u1._w_cache.get(u1.mlp.weight), u1.mlp.bias, u1.norm.num_groups,
                    sc._w_cache.get(sc.mlp.weight), sc.mlp.bias, sc.norm.num_groups, seg, 6)
                y = u1.apply(x, st1, seg, 6, 0.1).view(nq, 6, u1.out_dim)
            else:
                y = self.unary1(x, seg=s_seg)
        else:
            y = x
        y = _conv_double_norm(self.interso3, self.norm, y, q_pts, s_pts, neighb_inds, seg)
        if 'strided' in self.block_name:
            skip = K.maxpool_nbr(skip, neighb_inds.contiguous(), seg if sub_width is not None else None, sub_width)
        if nq > 0 and self.unary2.two_pass_ok() and (not has_skip_conv or self.skip_conv.two_pass_ok()):
            # statistics passes, then the Linear(s) recomputed with normalisation + residual + LeakyReLU in the
            # GEMM epilogue: the (N, A, out_dim) pre-norm tensors never reach global memory
            y3 = y.view(nq, 6, -1)
            _, st2 = self.unary2.pre_norm(y3, seg, 6, store=False)
            if has_skip_conv:
                st_s = st_skip if st_skip is not None else self.skip_conv.pre_norm(skip, seg, 6, store=False)[1]
                if _GFLAGS['dual_apply'] and dual_apply_supported(self.out_dim, y3.shape[-1], self.in_dim):
                    # both Linears recomputed into two accumulators of one kernel: the normalised shortcut is never stored
                    u2, sc = self.unary2, self.skip_conv
                    return linear_gn_apply_dual(
                        y3.reshape(-1, u2.in_dim), u2._w_cache.get(u2.mlp.weight), u2.mlp.bias, st2,
                        u2.norm.norm.weight, u2.norm.norm.bias,
                        _act(skip).reshape(-1, sc.in_dim), sc._w_cache.get(sc.mlp.weight), sc.mlp.bias, st_s,
                        sc.norm.norm.weight, sc.norm.norm.bias, u2.norm.norm.eps, 0.1, u2.norm.num_groups, seg,
                        6).view(nq, 6, self.out_dim)
                resid = self.skip_conv.apply(skip, st_s, seg, 6, 1.0)
            else:
                resid = skip.contiguous().view(nq * 6, self.out_dim)
            return self.unary2.apply(y3, st2, seg, 6, 0.1, resid=resid).view(nq, 6, self.out_dim)
        pre2, st2 = self.unary2.pre_norm(y.view(nq, 6, -1), seg, 6)
        if has_skip_conv:
            pre_s, st_s = self.skip_conv.pre_norm(skip, seg, 6)
            _, out = self.unary2.norm.fused(pre2, seg, 6, slope=0.1, other=(pre_s, self.skip_conv.norm, st_s),
                                            stats=st2)
        else:
            _, out = self.unary2.norm.fused(pre2, seg, 6, slope=0.1, resid=skip.contiguous(), stats=st2)
        return out.view(nq, 6, self.out_dim)


class InvOutBlockEPN(nn.Module):
    """blocks_epn.py:854-926 with att_pooling / att_permute off (the SE3ET configs): max over anchors."""

    def __init__(self, block_name, in_dim, config):
        super().__init__()
        if getattr(config, 'att_pooling', False) or getattr(config, 'att_permute', False):
            raise NotImplementedError("attentive pooling is not used by any SE3ET config")
        self.block_name, self.in_dim = block_name, in_dim

    def forward(self, x, q_pts=None, s_pts=None, neighb_inds=None):
        return K.anchor_max(_act(x).contiguous())


class LiftBlockEPN(nn.Module):
    """blocks_epn.py:993-1004: (N, C) -> (N, A, C)."""

    def __init__(self, block_name, in_dim, config):
        super().__init__()
        self.block_name, self.in_dim, self.kanchor = block_name, in_dim, config.kanchor

    def forward(self, x):
        return x.unsqueeze(1).expand(-1, self.kanchor, -1)


# ---- invariant decoder pieces (geotransformer/modules/kpconv/modules.py:33-101, functional.py:6-22) ----------


class GroupNorm(nn.Module):
    def __init__(self, num_groups, num_channels):
        super().__init__()
        self.num_groups, self.num_channels = num_groups, num_channels
        self.norm = nn.GroupNorm(self.num_groups, self.num_channels)

    def forward(self, x, seg=None):
        y = x.float().contiguous()
        seg = _seg(seg, x.shape[0], x.device)
        stats = K.groupnorm_stats(y, self.num_groups, seg, 1)
        _, out = K.groupnorm_apply(y, stats, self.norm.weight, self.norm.bias, self.num_groups, seg, 1, slope=1.0,
                                   eps=self.norm.eps)
        return out.squeeze()


class UnaryBlock(nn.Module):
    def __init__(self, in_channels, out_channels, group_norm, has_relu=True, bias=True, layer_norm=False):
        super().__init__()
        if layer_norm:
            raise NotImplementedError("layer_norm UnaryBlock is not used on the SE3ET path")
        self.in_channels, self.out_channels, self.group_norm = in_channels, out_channels, group_norm
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)
        self.norm = GroupNorm(group_norm, out_channels)
        self.leaky_relu = nn.LeakyReLU(0.1) if has_relu else None
        self._w_cache = _Bf16Cache()

    def forward(self, x, seg=None):
        seg = _seg(seg, x.shape[0], x.device)
        slope = 0.1 if self.leaky_relu is not None else 1.0
        xb, w = _act(x).contiguous(), self._w_cache.get(self.mlp.weight)
        if _GFLAGS['two_pass_unary'] and _gn_fusable(self.out_channels, self.norm.num_groups) and x.shape[0] > 0:
            _, stats = linear_gn_stats(xb, w, self.mlp.bias, self.norm.num_groups, seg, 1, store=False)
            return linear_gn_apply(xb, w, self.mlp.bias, stats, self.norm.norm.weight, self.norm.norm.bias,
                                   self.norm.norm.eps, slope, self.norm.num_groups, seg, 1)
        y, stats = linear_gn_stats(xb, w, self.mlp.bias, self.norm.num_groups, seg, 1)
        _, out = K.groupnorm_apply(y, stats, self.norm.norm.weight, self.norm.norm.bias, self.norm.num_groups, seg, 1,
                                   slope=0.1 if self.leaky_relu is not None else 1.0, eps=self.norm.norm.eps)
        return out


class LastUnaryBlock(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.mlp = nn.Linear(in_channels, out_channels, bias=bias)
        self._w_cache = _Bf16Cache()

    def forward(self, x, out_bf16=False):
        f, b = linear_bf16(_act(x).contiguous(), self._w_cache.get(self.mlp.weight), self.mlp.bias,
                           out_f32=not out_bf16, out_bf16=out_bf16)
        return b if out_bf16 else f


def nearest_upsample(x, upsample_indices):
    """kpconv/functional.py:6-22 (first neighbour column, zero shadow row)."""
    x = _act(x).contiguous()
    empty = torch.empty((upsample_indices.shape[0], 0), dtype=torch.bfloat16, device=x.device)
    return K.upsample_concat(x, upsample_indices, empty)


class E2PN(nn.Module):
    """Backbone of experiments/se3et*.3dmatch/backbone.py:8-77 (num_stages=4) and se3eti.kitti/backbone.py:8-99
    (num_stages=5).  forward(feats, data_dict) -> [feats_f, (latents...), feats_c]; feats_f is
    fp32 (N_1, output_dim), feats_c is the equivariant (N_last, A, 16*init_dim or 32*init_dim) tensor (bf16).
    data_dict may carry 'pair_offsets': per-level int64 offsets so that several pairs share one launch."""

    def __init__(self, input_dim, output_dim, init_dim, init_radius, init_sigma, group_norm, config_epn,
                 num_stages=4):
        super().__init__()
        R, S, G, c, d = init_radius, init_sigma, group_norm, config_epn, init_dim
        self.num_stages = num_stages
        self.preprocess = LiftBlockEPN('lift_epn', input_dim, c)
        self.encoder1_1 = SimpleBlockEPN('simple', input_dim, d, R, S, G, c)
        self.encoder1_2 = ResnetBottleneckBlockEPN('resnetb', d, d * 2, R, S, G, c)
        width = d * 2
        for s in range(2, num_stages + 1):
            lo, hi = 2 ** (s - 2), 2 ** (s - 1)
            setattr(self, 'encoder%d_1' % s, ResnetBottleneckBlockEPN('resnetb_strided', width, width, R * lo, S * lo, G, c))
            setattr(self, 'encoder%d_2' % s, ResnetBottleneckBlockEPN('resnetb', width, width * 2, R * hi, S * hi, G, c))
            setattr(self, 'encoder%d_3' % s, ResnetBottleneckBlockEPN('resnetb', width * 2, width * 2, R * hi, S * hi, G, c))
            setattr(self, 'equ2inv%d' % s, InvOutBlockEPN('inv_epn', width * 2, c))
            width *= 2
        for s in range(num_stages - 1, 2, -1):  # decoder{S-1} ... decoder3: UnaryBlock(24d*2^(s-3) -> 8d*2^(s-3))
            k = 2 ** (s - 3)
            setattr(self, 'decoder%d' % s, UnaryBlock(d * 24 * k, d * 8 * k, G))
        self.decoder2 = LastUnaryBlock(d * 12, output_dim)
        self.equ2inv = InvOutBlockEPN('inv_epn', output_dim, c)

    def forward(self, feats, data_dict):
        pts, nb = data_dict['points'], data_dict['neighbors']
        sub, up = data_dict['subsampling'], data_dict['upsampling']
        segs = data_dict.get('pair_offsets', [None] * len(pts))
        widths = data_dict.get('subsampling_width', [None] * len(pts))
        x = self.preprocess(_act(feats))  # stays an expand: the first conv recognises the lifted (anchor-constant) input
        x = self.encoder1_1(x, pts[0], pts[0], nb[0], seg=segs[0])
        x = self.encoder1_2(x, pts[0], pts[0], nb[0], seg=segs[0])
        inv = {}
        for s in range(2, self.num_stages + 1):
            l = s - 1
            x = getattr(self, 'encoder%d_1' % s)(x, pts[l], pts[l - 1], sub[l - 1], seg=segs[l], s_seg=segs[l - 1],
                                                 sub_width=widths[l - 1])
            x = getattr(self, 'encoder%d_2' % s)(x, pts[l], pts[l], nb[l], seg=segs[l])
            x = getattr(self, 'encoder%d_3' % s)(x, pts[l], pts[l], nb[l], seg=segs[l])
            inv[s] = K.anchor_max(x)
        feats_list = [x]
        latent = inv[self.num_stages]
        for s in range(self.num_stages - 1, 1, -1):
            cat = K.upsample_concat(latent, up[s - 1], inv[s])
            if s > 2:
                latent = getattr(self, 'decoder%d' % s)(cat, seg=segs[s - 1])
            else:
                latent = self.decoder2(cat)
            feats_list.append(latent)
        feats_list.reverse()
        return feats_list
