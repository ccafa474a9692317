"""Neighbour-limit calibration: GPU mirror of `calibrate_neighbors_stack_mode` (geotransformer/utils/data.py:212-252).

The reference walks its dataset, builds the pyramid of every pair with the histogram size as the neighbour limit,
histograms the neighbourhood sizes of every stage and keeps, per stage, the smallest limit that covers `keep_ratio` of
the points.  Same procedure here on point-cloud pairs handed in directly (no Dataset / collate layer: that is host
plumbing outside the hot path); the pyramid is built by se3et_b200.precompute on the GPU and only the per-stage
histograms travel to the host."""
import math

import numpy as np
import torch

from .ops import grid_subsample, radius_search


def neighbor_histograms(points, lengths, num_stages, voxel_size, search_radius, hist_n):
    """(num_stages, hist_n) int64: row s counts the points of stage s by their number of neighbours within the
    stage's search radius (the point itself included, as in the reference's self search), clipped to hist_n - 1
    columns the way np.bincount(c, minlength=hist_n)[:hist_n] drops larger counts."""
    dev = points.device
    lengths = lengths.to(dev)
    normals = torch.zeros_like(points)
    hists = torch.zeros((num_stages, hist_n), dtype=torch.int64)
    radius = search_radius
    for i in range(num_stages):
        if i > 0:
            points, lengths, normals = grid_subsample(points, lengths, normals, voxel_size=voxel_size)
        nb = radius_search(points, points, lengths, lengths, radius, hist_n)
        counts = (nb < points.shape[0]).sum(dim=1)
        h = torch.bincount(counts, minlength=hist_n)[:hist_n]
        hists[i] = h.cpu()
        voxel_size *= 2
        radius *= 2
    return hists.numpy()


def calibrate_neighbors_stack_mode(pairs, num_stages, voxel_size, search_radius, keep_ratio=0.8, sample_threshold=2000,
                                   device="cuda"):
    """pairs: iterable of (ref (n, 3), src (m, 3)) float32 arrays.  Returns int array (num_stages,) of neighbour
    limits, identical to the reference's for the same clouds (data.py:212-252)."""
    hist_n = int(math.ceil(4 / 3 * math.pi * (search_radius / voxel_size + 1) ** 3))
    neighbor_hists = np.zeros((num_stages, hist_n), dtype=np.int64)
    for ref, src in pairs:
        pts = torch.from_numpy(np.concatenate([ref, src]).astype(np.float32)).to(device)
        lens = torch.tensor([len(ref), len(src)], dtype=torch.int64)
        neighbor_hists += neighbor_histograms(pts, lens, num_stages, voxel_size, search_radius, hist_n)
        if np.min(np.sum(neighbor_hists, axis=1)) > sample_threshold:
            break
    cum_sum = np.cumsum(neighbor_hists.T, axis=0)
    return np.sum(cum_sum < (keep_ratio * cum_sum[hist_n - 1, :]), axis=0)
