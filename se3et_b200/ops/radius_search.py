"""Mirror of geotransformer/modules/ops/radius_search.py:7-27 on the CUDA path."""
from .. import _lib, ext


def radius_search(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit):
    """Neighbours of q_points in s_points within `radius` (stack mode), sorted by (distance, index), first
    `neighbor_limit` columns, padded with M = len(s_points). Shape (N, min(max_count, neighbor_limit)) exactly as
    the reference wrapper produces."""
    if neighbor_limit <= 0:
        return ext.radius_neighbors(q_points, s_points, q_lengths, s_lengths, radius)
    out, _, status = ext.radius_neighbors_raw(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit)
    max_count = int(status.cpu()[_lib.STATUS_MAX_COUNT])
    if max_count < neighbor_limit:
        out = out[:, :max_count].contiguous()
    return out


def radius_search_deferred(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit):
    """Same, without the host sync: returns the (N, neighbor_limit) matrix and the device status buffer whose
    word STATUS_MAX_COUNT the caller checks later (columns >= max_count are all padding)."""
    out, _, status = ext.radius_neighbors_raw(q_points, s_points, q_lengths, s_lengths, radius, neighbor_limit)
    return out, status
