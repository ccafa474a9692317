"""Drop-in for the reference's `geotransformer.ext` pybind module (extensions/pybind.cpp:6-18).

Same two functions, same argument order and dtypes, same error behaviour (RuntimeError on a
wrong dtype / non-contiguous tensor) -- except that tensors live on the GPU (the reference's
CHECK_CPU becomes CHECK_CUDA) and the two orders the reference leaves implementation-defined
are canonical (ascending voxel key; (d2, index) ascending).

    sys.modules['geotransformer.ext'] = se3et_b200.ext
"""
import ctypes
import threading

import torch

from . import _lib

DEFAULT_MAX_CELLS = 1 << 26  # voxel-bitmap budget (cells); grown automatically when a grid needs more


def _check(t, name, dtype, what):
    if not isinstance(t, torch.Tensor):
        raise RuntimeError("%s must be a tensor" % name)
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be a %s tensor" % (name, what))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def _lengths_on(lengths, name, device):
    if not isinstance(lengths, torch.Tensor) or lengths.dtype != torch.int64:
        raise RuntimeError("%s must be an long tensor" % name)
    if not lengths.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    return lengths.to(device, non_blocking=True)


def grid_subsampling_raw(points, lengths, normals, voxel_size, max_cells=None):
    """Launches the subsample; returns capacity-sized outputs, device lengths and the status buffer (no sync)."""
    _check(points, "points", torch.float32, "float")
    _check(normals, "normals", torch.float32, "float")
    lengths = _lengths_on(lengths, "lengths", points.device)
    n, b = points.shape[0], lengths.shape[0]
    if max_cells is None:
        max_cells = max(DEFAULT_MAX_CELLS, 64 * n)
    L = _lib.lib()
    nbytes = ctypes.c_size_t(0)
    _lib.check(L.se3et_grid_subsample_workspace_bytes(_lib.i64(n), _lib.i64(b), _lib.i64(max_cells),
                                                      ctypes.byref(nbytes)), "grid_subsample_workspace_bytes")
    ws = _lib.workspace.get(nbytes.value, points.device)
    s_points = torch.empty((max(n, 1), 3), dtype=torch.float32, device=points.device)
    s_normals = torch.empty((max(n, 1), 3), dtype=torch.float32, device=points.device)
    s_lengths = torch.empty((b,), dtype=torch.int64, device=points.device)
    status = torch.empty((_lib.SE3ET_STATUS_WORDS,), dtype=torch.int32, device=points.device)
    _lib.check(L.se3et_grid_subsample(
        _lib.ptr(points), _lib.ptr(lengths), _lib.ptr(normals), _lib.i64(n), _lib.i64(b), _lib.f32(voxel_size),
        _lib.ptr(s_points), _lib.ptr(s_lengths), _lib.ptr(s_normals), _lib.ptr(status), _lib.ptr(ws),
        ctypes.c_size_t(ws.numel()), _lib.i64(max_cells), _lib.stream_ptr()), "grid_subsample")
    return s_points, s_lengths, s_normals, status


# size of the largest cloud of the calling thread's last grid_subsampling (launch sequences run one per host thread)
class _Last(threading.local):
    def __init__(self):
        self.d = {}

    def __setitem__(self, k, v):
        self.d[k] = v

    def get(self, k, default=None):
        return self.d.get(k, default)


LAST = _Last()


def grid_subsampling(points, lengths, normals, voxel_size):
    """ext.grid_subsampling(points, lengths, normals, voxel_size) -> [s_points, s_lengths, s_normals]
    (grid_subsampling.h:6-11)."""
    max_cells = None
    for _ in range(3):
        s_points, s_lengths, s_normals, status = grid_subsampling_raw(points, lengths, normals, voxel_size, max_cells)
        st = status.cpu()  # the one sync: output size is data dependent
        err = int(st[_lib.STATUS_ERROR])
        if err & _lib.DEV_GRID_TOO_LARGE:
            max_cells = (int(st[_lib.STATUS_REQ_KCELLS]) + 1) * 1024
            if max_cells > (1 << 36):
                raise RuntimeError("grid_subsampling: voxel grid of %d Ki cells is too large" % (max_cells >> 10))
            continue
        if err:
            raise RuntimeError("grid_subsampling: device status %d" % err)
        m = int(st[_lib.STATUS_M_TOTAL])
        LAST['max_length'] = int(st[_lib.STATUS_MAX_LENGTH])   # rides on the same read-back (precompute's superpoint cap)
        return [s_points[:m], s_lengths, s_normals[:m]]
    raise RuntimeError("grid_subsampling: voxel grid does not fit the workspace")


def radius_neighbors_raw(q_points, s_points, q_lengths, s_lengths, radius, width, want_counts=False):
    """One launch sequence writing the first `width` sorted neighbours of every query (no sync).
    Returns (neighbors or None, counts or None, status); status = SE3ET_STATUS_WORDS words + per-cloud max counts."""
    _check(q_points, "q_points", torch.float32, "float")
    _check(s_points, "s_points", torch.float32, "float")
    dev = q_points.device
    q_lengths = _lengths_on(q_lengths, "q_lengths", dev)
    s_lengths = _lengths_on(s_lengths, "s_lengths", dev)
    nq, ns, b = q_points.shape[0], s_points.shape[0], q_lengths.shape[0]
    L = _lib.lib()
    nbytes = ctypes.c_size_t(0)
    _lib.check(L.se3et_radius_neighbors_workspace_bytes(_lib.i64(nq), _lib.i64(ns), _lib.i64(b), ctypes.byref(nbytes)),
               "radius_neighbors_workspace_bytes")
    ws = _lib.workspace.get(nbytes.value, dev)
    out = torch.empty((nq, width), dtype=torch.int64, device=dev) if width > 0 else None
    counts = torch.empty((max(nq, 1),), dtype=torch.int32, device=dev) if want_counts else None
    # status words followed by the per-cloud maximum counts (one small buffer, read together)
    status = torch.empty((_lib.SE3ET_STATUS_WORDS + b,), dtype=torch.int32, device=dev)
    cloud_max = status[_lib.SE3ET_STATUS_WORDS:]
    _lib.check(L.se3et_radius_neighbors(
        _lib.ptr(q_points), _lib.ptr(s_points), _lib.ptr(q_lengths), _lib.ptr(s_lengths), _lib.i64(nq), _lib.i64(ns),
        _lib.i64(b), _lib.f32(radius), _lib.ptr(counts), _lib.ptr(out), _lib.i64(width), _lib.ptr(status),
        _lib.ptr(cloud_max), _lib.ptr(ws), ctypes.c_size_t(ws.numel()), _lib.stream_ptr()), "radius_neighbors")
    return out, counts, status


def radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius):
    """ext.radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius) -> (Nq, max_count) int64,
    rows padded with Ns (radius_neighbors.h:5-11). Two passes: count, then fill at the exact width."""
    _, _, status = radius_neighbors_raw(q_points, s_points, q_lengths, s_lengths, radius, 0)
    width = int(status.cpu()[_lib.STATUS_MAX_COUNT])
    if width == 0 or q_points.shape[0] == 0:
        return torch.zeros((q_points.shape[0], width), dtype=torch.int64, device=q_points.device)
    out, _, _ = radius_neighbors_raw(q_points, s_points, q_lengths, s_lengths, radius, width)
    return out
