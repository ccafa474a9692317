"""GPU: se3et_point_to_node_partition (through the C ABI) against the numpy oracle, bit-exact on every index output."""
import os

import numpy as np
import pytest
import torch

from oracle import partition as opart

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "partition_ref.npz")


def run_stacked(pts, pl, nodes, nl, k, return_count=True):
    from se3et_b200.ops.partition_ops import point_to_node_partition_stacked
    out = point_to_node_partition_stacked(torch.from_numpy(pts).to(DEV), torch.tensor(pl, dtype=torch.int64, device=DEV),
                                          torch.from_numpy(nodes).to(DEV), torch.tensor(nl, dtype=torch.int64, device=DEV),
                                          k, return_count=return_count)
    return [o.cpu().numpy() for o in out]


def check(pts, pl, nodes, nl, k):
    p2n, masks, knn, knn_masks, sizes = run_stacked(pts, pl, nodes, nl, k)
    w_p2n, w_sizes, w_masks, w_knn, w_km = opart.point_to_node_partition_stacked(pts, pl, nodes, nl, k)
    assert np.array_equal(p2n, w_p2n)
    assert np.array_equal(sizes, w_sizes)
    assert np.array_equal(masks, w_masks)
    assert np.array_equal(knn_masks, w_km)
    assert np.array_equal(knn, w_knn)


def test_partition_matches_oracle_on_reference_fixtures():
    g = np.load(GOLD)
    for tag, k in (("demo", 64), ("demo8", 8), ("synth", 16)):
        pts = [g["%s_%d_points" % (tag, b)] for b in range(2)]
        nodes = [g["%s_%d_nodes" % (tag, b)] for b in range(2)]
        check(np.concatenate(pts), [len(p) for p in pts], np.concatenate(nodes), [len(n) for n in nodes], k)


def test_partition_reference_signature():
    """The drop-in of pointcloud_partition.point_to_node_partition: same argument and return order."""
    from se3et_b200.modules.partition import point_to_node_partition
    g = np.load(GOLD)
    pts, nodes = g["demo_0_points"], g["demo_0_nodes"]
    p2n, sizes, masks, knn, knn_masks = point_to_node_partition(torch.from_numpy(pts).to(DEV),
                                                                torch.from_numpy(nodes).to(DEV), 64, return_count=True)
    w = opart.point_to_node_partition(pts, nodes, 64)
    for got, want in zip((p2n, sizes, masks, knn, knn_masks), w):
        assert np.array_equal(got.cpu().numpy(), want)
    res = point_to_node_partition(torch.from_numpy(pts).to(DEV), torch.from_numpy(nodes).to(DEV), 64)
    assert len(res) == 4 and res[1].dtype == torch.bool and res[2].dtype == torch.int64


def test_partition_edge_cases():
    rng = np.random.default_rng(3)
    # clouds: regular, empty, one with duplicated nodes + an unreachable node, one with a single huge node (> 128
    # points: the exact slow path) and a limit larger than the cloud
    a = rng.random((700, 3), dtype=np.float32)
    c = rng.random((300, 3), dtype=np.float32)
    e = (rng.random((400, 3), dtype=np.float32) * 0.1).astype(np.float32)
    # grid points: many exactly equal distances (tie rule)
    gx = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(4), indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    pts = np.concatenate([a, c, e, gx])
    nodes_a = a[::9]
    nodes_c = np.concatenate([c[:5], c[:2], np.full((1, 3), 50.0, np.float32)])
    nodes_e = e[:1]
    nodes_g = np.array([[0, 0, 0], [7, 7, 3], [3.5, 3.5, 1.5], [3.5, 3.5, 1.5]], np.float32)
    nodes = np.concatenate([nodes_a, nodes_c, nodes_e, nodes_g])
    pl = [700, 0, 300, 400, len(gx)]
    nl = [len(nodes_a), 0, len(nodes_c), 1, 4]
    for k in (1, 7, 64, 500):
        check(pts, pl, nodes, nl, k)


def test_partition_full_size_properties():
    """3DMatch-sized stacked clouds (8 clouds of ~4.6k fine points, ~350 nodes): partition properties that do not need
    the oracle -- every point in exactly one node, rows sorted by distance, masks consistent with sizes."""
    rng = np.random.default_rng(5)
    pl = [4600 + 37 * i for i in range(8)]
    nl = [350 + 3 * i for i in range(8)]
    pts = np.concatenate([rng.random((n, 3), dtype=np.float32) * 3 for n in pl])
    po = np.concatenate([[0], np.cumsum(pl)])
    nodes = np.concatenate([pts[po[i]:po[i + 1]][::pl[i] // nl[i]][:nl[i]] for i in range(8)])
    p2n, masks, knn, knn_masks, sizes = run_stacked(pts, pl, nodes, nl, 64)
    no = np.concatenate([[0], np.cumsum(nl)])
    for b in range(8):
        P, N = pts[po[b]:po[b + 1]], nodes[no[b]:no[b + 1]]
        s = sizes[no[b]:no[b + 1]]
        assert s.sum() == pl[b] and np.array_equal(np.bincount(p2n[po[b]:po[b + 1]], minlength=nl[b]), s)
        assert np.array_equal(masks[no[b]:no[b + 1]], s > 0)
        km = knn_masks[no[b]:no[b + 1]]
        assert np.array_equal(km.sum(1), np.minimum(s, 64))
        kk = knn[no[b]:no[b + 1]]
        assert (kk[~km] == pl[b]).all()
        d = opart.sq_distances(N[:40], P)
        for j in range(40):
            sel = kk[j][km[j]]
            assert (p2n[po[b]:po[b + 1]][sel] == j).all() and (np.diff(d[j, sel]) >= 0).all()
