"""Host wrapper of se3et_point_to_node_partition (csrc/partition.cu)."""
import ctypes

import torch

from .. import _lib


def point_to_node_partition_stacked(points, point_lengths, nodes, node_lengths, point_limit, return_count=False):
    """Stacked clouds: points (N, 3) / nodes (M, 3) fp32 on the GPU, point_lengths / node_lengths int64 (B,) on the GPU.
    -> point_to_node (N,) int64, node_masks (M,) bool, node_knn_indices (M, K) int64, node_knn_masks (M, K) bool
    [, node_sizes (M,) int64]; all indices are cloud-local (pointcloud_partition.py:60-107 per cloud)."""
    _lib.require_cuda(points, nodes, point_lengths, node_lengths)
    assert points.dtype == torch.float32 and nodes.dtype == torch.float32
    assert point_lengths.dtype == torch.int64 and node_lengths.dtype == torch.int64
    assert point_lengths.numel() == node_lengths.numel() and point_lengths.numel() > 0
    points, nodes = points.contiguous(), nodes.contiguous()
    n, m, b, k = points.shape[0], nodes.shape[0], point_lengths.numel(), int(point_limit)
    dev = points.device
    p2n = torch.empty((n,), dtype=torch.int64, device=dev)
    masks = torch.empty((m,), dtype=torch.uint8, device=dev)
    sizes = torch.empty((m,), dtype=torch.int64, device=dev) if return_count else None
    knn = torch.empty((m, k), dtype=torch.int64, device=dev)
    knn_masks = torch.empty((m, k), dtype=torch.uint8, device=dev)
    nbytes = ctypes.c_size_t(0)
    _lib.check(_lib.lib().se3et_point_to_node_partition_workspace_bytes(_lib.i64(n), _lib.i64(b), ctypes.byref(nbytes)),
               "point_to_node_partition_workspace_bytes")
    ws = _lib.workspace.get(nbytes.value + 256, dev)
    off = (-ws.data_ptr()) % 256
    _lib.check(_lib.lib().se3et_point_to_node_partition(
        _lib.ptr(points), _lib.ptr(point_lengths), _lib.i64(n), _lib.ptr(nodes), _lib.ptr(node_lengths), _lib.i64(m),
        _lib.i64(b), _lib.i64(k), _lib.ptr(p2n), _lib.ptr(masks), _lib.ptr(sizes), _lib.ptr(knn), _lib.ptr(knn_masks),
        ctypes.c_void_p(ws.data_ptr() + off), ctypes.c_size_t(ws.numel() - off), _lib.stream_ptr()),
        "point_to_node_partition")
    out = (p2n, masks.bool(), knn, knn_masks.bool())
    return out + (sizes,) if return_count else out
