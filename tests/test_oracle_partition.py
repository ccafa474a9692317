"""CPU: the numpy restatement of point_to_node_partition (oracle/partition.py) against the outputs of the reference
function itself (tests/golden/partition_ref.npz, made by tests/golden/make_partition_golden.py)."""
import os

import numpy as np
import pytest

from oracle import partition as opart

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "partition_ref.npz")


def cases():
    g = np.load(GOLD)
    return g, sorted({k[:-len("points")] for k in g.files if k.endswith("_points")})


@pytest.mark.parametrize("key", cases()[1])
def test_oracle_matches_reference_partition(key):
    g = np.load(GOLD)
    pts, nodes, limit = g[key + "points"], g[key + "nodes"], int(g[key + "limit"])
    p2n, sizes, masks, knn, knn_masks = opart.point_to_node_partition(pts, nodes, limit)
    ref_p2n, ref_knn = g[key + "point_to_node"], g[key + "node_knn_indices"]
    d = opart.sq_distances(nodes, pts)
    # The reference's expanded formula x2 - 2 x.y + y2 cancels catastrophically (terms ~10, distances ~1e-3), and the
    # summation order of its matmul belongs to the BLAS: choices between distances that differ by less than that
    # rounding error can go either way (its own topk(k=8) and topk(k=64) already order such points differently).  The
    # restatement fixes the operation order; against the reference it may differ only inside that error bound.
    scale = float((nodes ** 2).sum(1).max() + (pts ** 2).sum(1).max())
    tol = 8 * np.finfo(np.float32).eps * scale
    # (1) assignment: identical except for near-ties
    diff = np.nonzero(p2n != ref_p2n)[0]
    assert len(diff) <= 0.01 * len(p2n)
    assert (np.abs(d[p2n[diff], diff] - d[ref_p2n[diff], diff]) <= tol).all()
    # (2) per-node selection, given the reference's own assignment
    sizes, masks, knn, knn_masks = opart.knn_from_assignment(d, ref_p2n, limit)
    assert np.array_equal(sizes, g[key + "node_sizes"])
    assert np.array_equal(masks, g[key + "node_masks"])
    assert np.array_equal(knn_masks, g[key + "node_knn_masks"])
    differing = np.nonzero((knn != ref_knn).any(1))[0]
    for j in differing:
        v = knn_masks[j]
        da, db = d[j, knn[j][v]], d[j, ref_knn[j][v]]
        assert (np.diff(db) >= -tol).all()                      # the reference's order is ascending within the bound
        assert np.abs(np.sort(da) - np.sort(db)).max() <= tol   # same distances (a boundary swap stays inside the bound)
    assert len(differing) <= 0.15 * len(knn)
    assert (knn[~knn_masks] == len(pts)).all() and (ref_knn[~knn_masks] == len(pts)).all()


def test_oracle_partition_edge_cases():
    rng = np.random.default_rng(0)
    pts = rng.random((50, 3), dtype=np.float32)
    # a node nobody is closest to (far away), duplicated nodes (tie -> lowest index), limit smaller than a node's size
    nodes = np.concatenate([pts[:3], pts[:1], np.full((1, 3), 100.0, np.float32)])
    p2n, sizes, masks, knn, knn_masks = opart.point_to_node_partition(pts, nodes, 4)
    assert p2n[0] == 0 and sizes[3] == 0 and not masks[3] and not masks[4]
    assert (knn[3] == 50).all() and not knn_masks[3].any()
    assert sizes.sum() == 50 and (knn_masks.sum(1) == np.minimum(sizes, 4)).all()
    d = opart.sq_distances(nodes, pts)
    for j in range(3):
        sel = knn[j][knn_masks[j]]
        assert (np.diff(d[j, sel]) >= 0).all() and (p2n[sel] == j).all()
    # stacked form: cloud-local indices
    st = opart.point_to_node_partition_stacked(np.concatenate([pts, pts[:20]]), [50, 20],
                                               np.concatenate([nodes, nodes[:2]]), [5, 2], 4)
    assert np.array_equal(st[0][:50], p2n) and st[0][50:].max() <= 1 and st[3].shape == (7, 4)
    assert (st[3][5:][~st[4][5:]] == 20).all()
