"""TEST INFRASTRUCTURE ONLY -- the product path (se3et_b200/) never imports this.

ctypes front-ends for
  * oracle/liboracle_points.so   our C restatement (oracle/points.c), and
  * oracle/_ref/libse3et_ref.so  the unmodified reference CPU path (oracle/ref_shim.cpp
                                 + the reference's own .cpp files, built by oracle/Makefile),
plus the canonicalisation wrappers SURVEY.md 8(c) requires where the reference's
result order is implementation-defined, and a numpy restatement of
`precompute_data_stack_mode` (geotransformer/utils/data.py:13-97) on top of them.

Allowed importers: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
--impl reference legs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(_HERE, "liboracle_points.so")
_REF_SO = os.path.join(_HERE, "_ref", "libse3et_ref.so")

_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    """Compile the C restatement and (when /root/reference is present) oracle/_ref."""
    if force or not os.path.exists(_ORACLE_SO):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle_points.so"])
    if os.path.isdir("/root/reference/geotransformer/extensions") and (force or not os.path.exists(_REF_SO)):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def _ptr(a, t):
    return a.ctypes.data_as(t)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


_libs = {}


def _oracle_lib():
    if "o" not in _libs:
        build()
        lib = ctypes.CDLL(_ORACLE_SO)
        lib.oracle_grid_subsample.restype = ctypes.c_int64
        lib.oracle_grid_subsample.argtypes = [_f32p, _i64p, _f32p, ctypes.c_int64, ctypes.c_float, _f32p, _i64p, _f32p]
        lib.oracle_radius_neighbors.restype = ctypes.c_int64
        lib.oracle_radius_neighbors.argtypes = [
            _f32p, _f32p, _i64p, _i64p, ctypes.c_int64, ctypes.c_float, _i64p, _i64p, ctypes.c_int64,
        ]
        _libs["o"] = lib
    return _libs["o"]


def have_ref():
    return os.path.exists(_REF_SO)


def _ref_lib():
    if "r" not in _libs:
        build()
        lib = ctypes.CDLL(_REF_SO)
        lib.ref_grid_subsampling.restype = ctypes.c_long
        lib.ref_grid_subsampling.argtypes = [
            _f32p, _i64p, _f32p, ctypes.c_long, ctypes.c_long, ctypes.c_float, _f32p, _i64p, _f32p, ctypes.c_long,
        ]
        lib.ref_radius_neighbors.restype = ctypes.c_long
        lib.ref_radius_neighbors.argtypes = [
            _f32p, _f32p, _i64p, _i64p, ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_float, _i64p,
            ctypes.c_long,
        ]
        _libs["r"] = lib
    return _libs["r"]


# ----------------------------------------------------------------------------------------------
# oracle (C restatement)
# ----------------------------------------------------------------------------------------------


def grid_subsample(points, lengths, normals, voxel_size):
    """Canonical-order restatement of ext.grid_subsampling. Returns (s_points, s_lengths, s_normals)."""
    points, normals, lengths = _f32(points), _f32(normals), _i64(lengths)
    n = points.shape[0]
    sp = np.empty((max(n, 1), 3), np.float32)
    sn = np.empty((max(n, 1), 3), np.float32)
    sl = np.zeros(lengths.shape[0], np.int64)
    m = _oracle_lib().oracle_grid_subsample(
        _ptr(points, _f32p), _ptr(lengths, _i64p), _ptr(normals, _f32p), lengths.shape[0],
        ctypes.c_float(voxel_size), _ptr(sp, _f32p), _ptr(sl, _i64p), _ptr(sn, _f32p),
    )
    return sp[:m].copy(), sl, sn[:m].copy()


def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius, return_counts=False):
    """Canonical-order restatement of ext.radius_neighbors: (Nq, max_count) int64, padded with Ns."""
    q, s = _f32(q_points), _f32(s_points)
    ql, sl = _i64(q_lengths), _i64(s_lengths)
    nq = q.shape[0]
    counts = np.zeros(max(nq, 1), np.int64)
    lib = _oracle_lib()
    width = lib.oracle_radius_neighbors(
        _ptr(q, _f32p), _ptr(s, _f32p), _ptr(ql, _i64p), _ptr(sl, _i64p), ql.shape[0], ctypes.c_float(radius),
        _ptr(counts, _i64p), None, 0,
    )
    out = np.empty((nq, width), np.int64)
    if nq and width:
        lib.oracle_radius_neighbors(
            _ptr(q, _f32p), _ptr(s, _f32p), _ptr(ql, _i64p), _ptr(sl, _i64p), ql.shape[0], ctypes.c_float(radius),
            None, _ptr(out, _i64p), width,
        )
    return (out, counts[:nq]) if return_counts else out


def radius_search(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit):
    """modules/ops/radius_search.py:7-27 on the oracle."""
    nb = radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius)
    if neighbor_limit > 0:
        nb = nb[:, :neighbor_limit]
    return np.ascontiguousarray(nb)


# ----------------------------------------------------------------------------------------------
# reference (unmodified C++), raw and canonicalised
# ----------------------------------------------------------------------------------------------


def ref_grid_subsampling_raw(points, lengths, normals, voxel_size):
    points, normals, lengths = _f32(points), _f32(normals), _i64(lengths)
    n = points.shape[0]
    sp = np.empty((max(n, 1), 3), np.float32)
    sn = np.empty((max(n, 1), 3), np.float32)
    sl = np.zeros(lengths.shape[0], np.int64)
    m = _ref_lib().ref_grid_subsampling(
        _ptr(points, _f32p), _ptr(lengths, _i64p), _ptr(normals, _f32p), n, lengths.shape[0],
        ctypes.c_float(voxel_size), _ptr(sp, _f32p), _ptr(sl, _i64p), _ptr(sn, _f32p), max(n, 1),
    )
    assert m >= 0
    return sp[:m].copy(), sl, sn[:m].copy()


def ref_radius_neighbors_raw(q_points, s_points, q_lengths, s_lengths, radius):
    q, s = _f32(q_points), _f32(s_points)
    ql, sl = _i64(q_lengths), _i64(s_lengths)
    lib = _ref_lib()
    args = (_ptr(q, _f32p), _ptr(s, _f32p), _ptr(ql, _i64p), _ptr(sl, _i64p), q.shape[0], s.shape[0], ql.shape[0],
            ctypes.c_float(radius))
    width = lib.ref_radius_neighbors(*args, None, 0)
    out = np.empty((q.shape[0], width), np.int64)
    if out.size:
        assert lib.ref_radius_neighbors(*args, _ptr(out, _i64p), width) == width
    return out


def voxel_keys(points, voxel_size):
    """fp32 voxel keys of one cloud, same arithmetic as grid_subsampling_cpu.cpp:13-42 (numpy fp32 is IEEE,
    uncontracted)."""
    p = _f32(points)
    v = np.float32(voxel_size)
    inv = np.float32(1.0 / float(v))
    mn, mx = p.min(0), p.max(0)
    org = np.floor(mn * inv) * v
    nx = np.uint64(np.floor((mx[0] - org[0]) / v) + np.float32(1))
    ny = np.uint64(np.floor((mx[1] - org[1]) / v) + np.float32(1))
    ijk = np.floor((p - org) / v).astype(np.int64).astype(np.uint64)
    return ijk[:, 0] + nx * ijk[:, 1] + nx * ny * ijk[:, 2]


def canonicalize_subsample(points_in, lengths_in, voxel_size, s_points, s_lengths, s_normals):
    """Reorder a subsample result to ascending voxel key inside each cloud."""
    points_in = _f32(points_in)
    out_p, out_n = np.empty_like(s_points), np.empty_like(s_normals)
    a = c = 0
    for n_in, m in zip(np.asarray(lengths_in).tolist(), np.asarray(s_lengths).tolist()):
        cloud = points_in[a:a + n_in]
        # keys of the chosen points are recomputed against the *input* cloud's origin
        v = np.float32(voxel_size)
        inv = np.float32(1.0 / float(v))
        mn, mx = cloud.min(0), cloud.max(0)
        org = np.floor(mn * inv) * v
        nx = np.uint64(np.floor((mx[0] - org[0]) / v) + np.float32(1))
        ny = np.uint64(np.floor((mx[1] - org[1]) / v) + np.float32(1))
        ijk = np.floor((s_points[c:c + m] - org) / v).astype(np.int64).astype(np.uint64)
        keys = ijk[:, 0] + nx * ijk[:, 1] + nx * ny * ijk[:, 2]
        order = np.argsort(keys, kind="stable")
        ks = keys[order]  # uint64: a point just below the fp32 origin wraps to a huge key, as in the reference
        assert np.all(ks[1:] > ks[:-1]), "two output points share a voxel"
        out_p[c:c + m] = s_points[c:c + m][order]
        out_n[c:c + m] = s_normals[c:c + m][order]
        a += n_in
        c += m
    return out_p, np.asarray(s_lengths).copy(), out_n


def canonicalize_neighbors(q_points, s_points, neighbors):
    """Stable re-sort of every row by (d2 asc, index asc). Padding (== Ns) stays last."""
    q, s = _f32(q_points), _f32(s_points)
    ns = s.shape[0]
    if neighbors.size == 0:
        return neighbors.copy()
    spad = np.concatenate([s, np.full((1, 3), np.inf, np.float32)], 0)
    nb = spad[neighbors]  # (Nq, W, 3)
    with np.errstate(invalid="ignore"):
        dx = q[:, None, 0] - nb[:, :, 0]
        dy = q[:, None, 1] - nb[:, :, 1]
        dz = q[:, None, 2] - nb[:, :, 2]
        d2 = (np.float32(0) + dx * dx) + dy * dy
        d2 = d2 + dz * dz
    d2[neighbors == ns] = np.inf
    order = np.lexsort((neighbors, d2), axis=-1)
    return np.take_along_axis(neighbors, order, axis=-1)


def ref_grid_subsample(points, lengths, normals, voxel_size):
    sp, sl, sn = ref_grid_subsampling_raw(points, lengths, normals, voxel_size)
    return canonicalize_subsample(points, lengths, voxel_size, sp, sl, sn)


def ref_radius_search(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit):
    nb = canonicalize_neighbors(q_points, s_points,
                                ref_radius_neighbors_raw(q_points, s_points, q_lengths, s_lengths, radius))
    if neighbor_limit > 0:
        nb = nb[:, :neighbor_limit]
    return np.ascontiguousarray(nb)


# ----------------------------------------------------------------------------------------------
# pyramid precompute (utils/data.py:13-97), parameterised by the two ops
# ----------------------------------------------------------------------------------------------


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits, impl="oracle"):
    """Restates geotransformer/utils/data.py:13-97 (normals are zeros: open3d normals are never
    consumed by the model). impl: 'oracle' | 'ref' (canonicalised) | 'ref_raw'."""
    assert num_stages == len(neighbor_limits)
    if impl == "oracle":
        sub, search = grid_subsample, radius_search
    elif impl == "ref":
        sub, search = ref_grid_subsample, ref_radius_search
    else:
        def sub(p, l, n, v):
            return ref_grid_subsampling_raw(p, l, n, v)

        def search(q, s, ql, sl, r, lim):
            nb = ref_radius_neighbors_raw(q, s, ql, sl, r)
            return np.ascontiguousarray(nb[:, :lim] if lim > 0 else nb)

    points = _f32(points)
    lengths = _i64(lengths).copy()
    normals = np.zeros_like(points)
    points_list, lengths_list, normals_list = [], [], []
    for i in range(num_stages):
        if i > 0:
            points, lengths, normals = sub(points, lengths, normals, voxel_size)
        if i == num_stages - 1:  # data.py:34-43, cap of 2000 superpoints per cloud
            if lengths[0] > 2000:
                points = np.concatenate([points[:2000], points[lengths[0]:]], 0)
                normals = np.concatenate([normals[:2000], normals[lengths[0]:]], 0)
                lengths[0] = 2000
            if lengths[1] > 2000:
                points = np.concatenate([points[:lengths[0]], points[lengths[0]:lengths[0] + 2000]], 0)
                normals = np.concatenate([normals[:lengths[0]], normals[lengths[0]:lengths[0] + 2000]], 0)
                lengths[1] = 2000
        points_list.append(points)
        lengths_list.append(lengths)
        normals_list.append(normals)
        voxel_size *= 2
    neighbors_list, subsampling_list, upsampling_list = [], [], []
    for i in range(num_stages):
        cp, cl = points_list[i], lengths_list[i]
        neighbors_list.append(search(cp, cp, cl, cl, radius, neighbor_limits[i]))
        if i < num_stages - 1:
            sp_, sl_ = points_list[i + 1], lengths_list[i + 1]
            subsampling_list.append(search(sp_, cp, sl_, cl, radius, neighbor_limits[i]))
            upsampling_list.append(search(cp, sp_, cl, sl_, radius * 2, neighbor_limits[i + 1]))
        radius *= 2
    return {
        "points": points_list, "lengths": lengths_list, "neighbors": neighbors_list,
        "subsampling": subsampling_list, "upsampling": upsampling_list, "normals": normals_list,
    }
