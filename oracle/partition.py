"""TEST ORACLE (not a product path): numpy restatement of point_to_node_partition
(geotransformer/modules/ops/pointcloud_partition.py:60-107) with the distance of pairwise_distance
(geotransformer/modules/ops/pairwise_distance.py:19-31).

Pinned by tests/test_oracle_partition.py against tests/golden/partition_ref.npz, the output of the reference function itself
(tests/golden/make_partition_golden.py runs it on the CPU with `.cuda()` patched out).  The reference leaves the
summation order of its matmul and the order of tied distances to the BLAS / torch.topk; this restatement fixes
  xy = (x0 y0 + x1 y1) + x2 y2,  d = (|x|^2 - 2 xy) + |y|^2 clamped at 0, every operation rounded to fp32 (no FMA),
  argmin -> lowest node index, top-k -> ascending (distance, point index),
which is what the CUDA kernel (se3et_b200/csrc/partition.cu) computes.
"""
import numpy as np


def sq_distances(nodes, points):
    """(M, N) fp32, pairwise_distance.py:27-30 with x = nodes, y = points."""
    x = np.ascontiguousarray(nodes, dtype=np.float32)
    y = np.ascontiguousarray(points, dtype=np.float32)
    x2 = (x[:, 0] * x[:, 0] + x[:, 1] * x[:, 1]) + x[:, 2] * x[:, 2]
    y2 = (y[:, 0] * y[:, 0] + y[:, 1] * y[:, 1]) + y[:, 2] * y[:, 2]
    xy = (x[:, None, 0] * y[None, :, 0] + x[:, None, 1] * y[None, :, 1]) + x[:, None, 2] * y[None, :, 2]
    d = (x2[:, None] - np.float32(2.0) * xy) + y2[None, :]
    return np.maximum(d, np.float32(0.0)).astype(np.float32)


def knn_from_assignment(d, p2n, point_limit):
    """Per node: its assigned points in ascending (distance, index) order, first point_limit (:88-97).
    d (M, N) distances, p2n (N,) -> node_sizes, node_masks, node_knn_indices (padded with N), node_knn_masks."""
    m, n = d.shape
    k = int(point_limit)
    sizes = np.bincount(p2n, minlength=m).astype(np.int64)[:m] if n > 0 else np.zeros((m,), np.int64)
    masks = sizes > 0                                                                     # (:85-86)
    knn = np.full((m, k), n, dtype=np.int64)                                              # (:97)
    knn_masks = np.zeros((m, k), dtype=bool)
    for j in range(m):
        idx = np.nonzero(p2n == j)[0]
        order = np.lexsort((idx, d[j, idx]))[:k]                                           # topk of the masked row (:93)
        knn[j, :len(order)] = idx[order]
        knn_masks[j, :len(order)] = True
    return sizes, masks, knn, knn_masks


def point_to_node_partition(points, nodes, point_limit):
    """-> point_to_node (N,) int64, node_sizes (M,) int64, node_masks (M,) bool, node_knn_indices (M, K) int64 (padded
    with N), node_knn_masks (M, K) bool.  pointcloud_partition.py:82-107."""
    n, m = points.shape[0], nodes.shape[0]
    d = sq_distances(nodes, points)
    p2n = np.argmin(d, axis=0).astype(np.int64) if m > 0 else np.zeros((n,), np.int64)  # first minimum (:84)
    return (p2n,) + knn_from_assignment(d, p2n, point_limit)


def point_to_node_partition_stacked(points, point_lengths, nodes, node_lengths, point_limit):
    """Per-cloud application; indices cloud-local, outputs concatenated."""
    outs = [[], [], [], [], []]
    po = no = 0
    for pl, nl in zip(point_lengths, node_lengths):
        res = point_to_node_partition(points[po:po + pl], nodes[no:no + nl], point_limit)
        for o, r in zip(outs, res):
            o.append(r)
        po += pl
        no += nl
    return tuple(np.concatenate(o) for o in outs)
