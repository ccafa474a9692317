"""Mirror of geotransformer/modules/ops/grid_subsample.py:9-25 on the CUDA path."""
from .. import ext


def grid_subsample(points, lengths, normals, voxel_size):
    """Grid subsampling in stack mode (GPU).

    Args / returns as the reference wrapper: (N,3) points, (B,) lengths, (N,3) normals, voxel size ->
    s_points (M,3), s_lengths (B,), s_normals (M,3). Output order: ascending voxel key per cloud.
    """
    s_points, s_lengths, s_normals = ext.grid_subsampling(points, lengths, normals, voxel_size)
    return s_points, s_lengths, s_normals
