"""GPU parity: CUDA grid_subsample / radius_neighbors (through the C ABI) vs the oracle and the committed
reference fixtures. Bit-exact: points, lengths and neighbour indices must be identical."""
import os

import numpy as np
import pytest
import torch

from oracle import points as op
from se3et_b200 import _lib, ext, synthetic
from se3et_b200.ops import grid_subsample, radius_search
from se3et_b200.precompute import precompute_data_stack_mode

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(autouse=True, params=["per_query", "by_cell"])
def radius_mode(request):
    """Every test of this file runs with both neighbour-search kernels: one warp per query (mode 0) and the by-cell
    kernel (mode 1; automatic selection only picks it for large self searches, which the small fixtures are not)."""
    L = _lib.lib()
    assert L.se3et_radius_set_mode(0 if request.param == "per_query" else 1) == 0
    yield request.param
    assert L.se3et_radius_set_mode(2) == 0


def _t(a, dtype=None):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def test_native_library_is_the_path():
    assert os.path.exists(_lib.lib()._name)


@pytest.mark.parametrize("name,stages", [("points_demo_crop.npz", 4), ("points_synth_small.npz", 3)])
def test_pyramid_matches_reference_fixture(golden_dir, name, stages):
    g = np.load(os.path.join(golden_dir, name))
    d = precompute_data_stack_mode(_t(g["in_points"]), _t(g["in_lengths"]), stages, float(g["voxel"]),
                                   float(g["radius"]), g["limits"].tolist())
    for i in range(stages):
        assert np.array_equal(g["points_%d" % i], d["points"][i].cpu().numpy())
        assert np.array_equal(g["lengths_%d" % i], d["lengths"][i].cpu().numpy())
        assert np.array_equal(g["neighbors_%d" % i], d["neighbors"][i].cpu().numpy())
        if i < stages - 1:
            assert np.array_equal(g["subsampling_%d" % i], d["subsampling"][i].cpu().numpy())
            assert np.array_equal(g["upsampling_%d" % i], d["upsampling"][i].cpu().numpy())


def _edge_clouds():
    rng = np.random.default_rng(0)
    cases = {}
    a = rng.uniform(-1, 1, (500, 3)).astype(np.float32)
    cases["ragged"] = (np.concatenate([a, a[:1] + 5, rng.uniform(0, 0.3, (37, 3)).astype(np.float32)]),
                       np.array([500, 1, 37]))
    lat = np.stack(np.meshgrid(np.arange(8), np.arange(8), np.arange(4), indexing="ij"), -1).reshape(-1, 3)
    lat = (lat * 0.03125).astype(np.float32)
    cases["lattice_ties"] = (np.concatenate([lat, lat[::3]]), np.array([len(lat), len(lat[::3])]))
    b = (rng.integers(-40, 40, (800, 3)) * 0.0125).astype(np.float32) + np.float32(0.65)
    cases["borders"] = (b, np.array([300, 500]))
    cases["empty_cloud"] = (a[:100], np.array([60, 0, 40]))
    return cases


@pytest.mark.parametrize("case", ["ragged", "lattice_ties", "borders", "empty_cloud"])
def test_edge_cases_match_oracle(case):
    pts, lens = _edge_clouds()[case]
    nrm = np.random.default_rng(1).normal(size=pts.shape).astype(np.float32)
    for voxel in (0.05, 0.1):
        o = op.grid_subsample(pts, lens, nrm, voxel)
        c = grid_subsample(_t(pts), _t(lens), _t(nrm), voxel)
        for x, y in zip(o, c):
            assert np.array_equal(x, y.cpu().numpy())
    for radius in (0.0625, 0.11):
        o = op.radius_neighbors(pts, pts, lens, lens, radius)
        c = ext.radius_neighbors(_t(pts), _t(pts), _t(lens), _t(lens), radius)
        assert np.array_equal(o, c.cpu().numpy())
        o = op.radius_search(pts, pts, lens, lens, radius, 5)
        c = radius_search(_t(pts), _t(pts), _t(lens), _t(lens), radius, 5)
        assert np.array_equal(o, c.cpu().numpy())


def test_dense_neighbourhoods_take_the_slow_exact_path():
    # > 256 neighbours per query: exceeds the shared-memory hit buffer
    rng = np.random.default_rng(5)
    pts = rng.uniform(0, 0.2, (1500, 3)).astype(np.float32)
    lens = np.array([1500])
    o = op.radius_search(pts, pts, lens, lens, 0.15, 300)
    c = radius_search(_t(pts), _t(pts), _t(lens), _t(lens), 0.15, 300)
    assert (o < 1500).sum(1).max() > 256
    assert np.array_equal(o, c.cpu().numpy())


def test_query_and_support_differ():
    rng = np.random.default_rng(6)
    s = rng.uniform(0, 1, (3000, 3)).astype(np.float32)
    q = rng.uniform(-0.2, 1.2, (700, 3)).astype(np.float32)
    ql, sl = np.array([300, 400]), np.array([1000, 2000])
    o = op.radius_neighbors(q, s, ql, sl, 0.1)
    c = ext.radius_neighbors(_t(q), _t(s), _t(ql), _t(sl), 0.1)
    assert np.array_equal(o, c.cpu().numpy())


def test_full_size_synthetic_pair_matches_oracle():
    p = synthetic.make_3dmatch_pair(0)
    pts = np.concatenate([p["ref_points"], p["src_points"]])
    lens = np.array([len(p["ref_points"]), len(p["src_points"])])
    o = op.precompute_data_stack_mode(pts, lens, 4, 0.025, 0.0625, [38, 36, 36, 38], impl="oracle")
    d = precompute_data_stack_mode(_t(pts), _t(lens), 4, 0.025, 0.0625, [38, 36, 36, 38])
    for k in ("points", "lengths", "neighbors", "subsampling", "upsampling"):
        for x, y in zip(o[k], d[k]):
            assert np.array_equal(x, y.cpu().numpy()), k


def test_batch_of_pairs_equals_pair_by_pair():
    # size-independent property at batch scale: stacking 8 pairs (16 clouds) gives the per-pair results
    pairs = [synthetic.make_3dmatch_pair(i, target_points=5000) for i in range(8)]
    clouds = [c for p in pairs for c in (p["ref_points"], p["src_points"])]
    pts = np.concatenate(clouds)
    lens = np.array([len(c) for c in clouds])
    sp, sl, _ = grid_subsample(_t(pts), _t(lens), _t(np.zeros_like(pts)), 0.05)
    nb = radius_search(_t(pts), _t(pts), _t(lens), _t(lens), 0.0625, 38).cpu().numpy()
    sp, sl = sp.cpu().numpy(), sl.cpu().numpy()
    a = c = 0
    for i, cl in enumerate(clouds):
        one_p, one_l, _ = grid_subsample(_t(cl), _t(np.array([len(cl)])), _t(np.zeros_like(cl)), 0.05)
        assert int(one_l[0]) == sl[i]
        assert np.array_equal(one_p.cpu().numpy(), sp[c:c + sl[i]])
        one_nb = radius_search(_t(cl), _t(cl), _t(np.array([len(cl)])), _t(np.array([len(cl)])), 0.0625, 38).cpu().numpy()
        batch_nb = nb[a:a + len(cl)]
        w = one_nb.shape[1]
        pad_one, pad_b = one_nb == len(cl), batch_nb[:, :w] == len(pts)
        assert np.array_equal(pad_one, pad_b)
        assert np.array_equal(np.where(pad_one, 0, one_nb + a), np.where(pad_b, 0, batch_nb[:, :w]))
        assert np.all(batch_nb[:, w:] == len(pts))
        a += len(cl)
        c += sl[i]


def test_superpoint_cap_applies_to_stacked_pairs_too():
    """data.py:34-43 keeps at most 2000 superpoints per cloud; stacked launch sequences apply the same cap per cloud,
    so that a pair gives the same pyramid batched and alone."""
    rng = np.random.default_rng(7)
    clouds = [rng.uniform(0, 1, (n, 3)).astype(np.float32) for n in (6000, 3000, 1500, 5000)]
    pts, lens = np.concatenate(clouds), np.array([len(c) for c in clouds])
    d = precompute_data_stack_mode(_t(pts), _t(lens), 2, 0.005, 0.0125, [8, 8])
    got_l = d["lengths"][1].cpu().numpy()
    assert got_l.max() == 2000 and got_l[2] < 2000
    off = np.concatenate([[0], np.cumsum(got_l)])
    for p in range(2):
        one = precompute_data_stack_mode(_t(np.concatenate(clouds[2 * p:2 * p + 2])), _t(lens[2 * p:2 * p + 2]), 2, 0.005,
                                         0.0125, [8, 8])
        assert np.array_equal(one["lengths"][1].cpu().numpy(), got_l[2 * p:2 * p + 2])
        assert np.array_equal(one["points"][1].cpu().numpy(), d["points"][1][off[2 * p]:off[2 * p + 2]].cpu().numpy())


def test_rejects_bad_arguments():
    pts = torch.zeros(4, 3, device=DEV)
    with pytest.raises(RuntimeError, match="float"):
        ext.grid_subsampling(pts.double(), torch.tensor([4]), pts, 0.1)
    with pytest.raises(RuntimeError, match="contiguous"):
        ext.radius_neighbors(torch.zeros(3, 4, device=DEV).t(), pts, torch.tensor([4]), torch.tensor([4]), 0.1)


@pytest.mark.parametrize("name", ["tdm_small", "kitti_small"])
def test_neighbor_limit_calibration_matches_oracle_and_reference(golden_dir, name):
    """se3et_b200.calibrate (GPU pyramid + histograms) == oracle == the unmodified reference function's limits."""
    import os
    from oracle import calibrate as ocal
    from se3et_b200.calibrate import calibrate_neighbors_stack_mode
    from test_oracle_calibrate import clouds_of
    gold = np.load(os.path.join(golden_dir, "calibrate_ref.npz"))
    stages, voxel, radius, keep, thresh = gold[name + "_params"]
    clouds = clouds_of(name)
    got = calibrate_neighbors_stack_mode(clouds, int(stages), float(voxel), float(radius), float(keep), int(thresh))
    assert np.array_equal(got, gold[name + "_limits"]), (got, gold[name + "_limits"])
    assert np.array_equal(got, ocal.calibrate_neighbors_stack_mode(clouds, int(stages), float(voxel), float(radius),
                                                                   float(keep), int(thresh)))
