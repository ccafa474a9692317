"""TEST INFRASTRUCTURE (never imported by se3et_b200/): numpy restatement of `calibrate_neighbors_stack_mode`
(geotransformer/utils/data.py:212-252) on the oracle pyramid primitives (oracle/points.py).  Pinned against the
unmodified reference function by tests/golden/make_calibrate_golden.py -> tests/golden/calibrate_ref.npz."""
import math

import numpy as np

from . import points as op


def calibrate_neighbors_stack_mode(pairs, num_stages, voxel_size, search_radius, keep_ratio=0.8, sample_threshold=2000,
                                   impl="oracle"):
    """pairs: iterable of (ref (n, 3), src (m, 3)).  impl = 'oracle' | 'ref' (unmodified reference C++ operators)."""
    hist_n = int(math.ceil(4 / 3 * math.pi * (search_radius / voxel_size + 1) ** 3))
    hists = np.zeros((num_stages, hist_n), dtype=np.int64)
    sub = op.ref_grid_subsample if impl == "ref" else op.grid_subsample
    search = op.ref_radius_search if impl == "ref" else op.radius_search
    for ref, src in pairs:
        pts = np.concatenate([ref, src]).astype(np.float32)
        lens = np.array([len(ref), len(src)], dtype=np.int64)
        normals = np.zeros_like(pts)
        v, r = voxel_size, search_radius
        for i in range(num_stages):
            if i > 0:
                pts, lens, normals = sub(pts, lens, normals, v)
            nb = search(pts, pts, lens, lens, r, hist_n)
            counts = np.sum(nb < nb.shape[0], axis=1)
            hists[i] += np.bincount(counts, minlength=hist_n)[:hist_n]
            v *= 2
            r *= 2
        if np.min(np.sum(hists, axis=1)) > sample_threshold:
            break
    cum = np.cumsum(hists.T, axis=0)
    return np.sum(cum < (keep_ratio * cum[hist_n - 1, :]), axis=0)
