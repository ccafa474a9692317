"""GPU pyramid precompute: mirror of geotransformer/utils/data.py:13-97 (`precompute_data_stack_mode`).

Same arguments and the same dict of lists, but every tensor stays on the GPU and the 3S-2 searches are issued
back-to-back without host syncs (their neighbour-count checks are read in one transfer at the end). Normals are
carried as zeros (the reference's open3d normals are never consumed by the model)."""
import torch

from . import _lib, ext
from .ops import grid_subsample, radius_search_deferred


def precompute_data_stack_mode(points, lengths, num_stages, voxel_size, radius, neighbor_limits, normals=None,
                               backbone_only=False):
    """`lengths` = [ref, src] reproduces the reference exactly (matrix widths min(max_count, limit), 2000-superpoint
    cap). `lengths` = [ref_1, src_1, ref_2, src_2, ...] stacks several pairs in one launch sequence: the dict then
    also carries 'pair_offsets' (per level, int64 [P+1]) and 'subsampling_width' (per level, int32 [P]: the width
    the reference's subsampling matrix would have for that pair alone), neighbour matrices keep `limit` columns
    (extra columns are padding) and no host sync is needed for the searches.
    backbone_only=True (what GeoTransformer.forward_stacked asks for): only what the E2PN backbone consumes is
    searched -- nearest_upsample reads column 0 of upsampling[1:] (kpconv/functional.py:21) and nothing reads
    upsampling[0] (the fine features live on level 1), so upsampling[0] is None and the others keep one column."""
    assert num_stages == len(neighbor_limits)
    num_pairs = lengths.shape[0] // 2
    batched = lengths.shape[0] > 2
    if batched:
        assert lengths.shape[0] % 2 == 0 and all(l > 0 for l in neighbor_limits)
    _lib.require_cuda(points)
    lengths = lengths.to(points.device)
    if normals is None:
        normals = torch.zeros_like(points)
    points_list, lengths_list, normals_list = [], [], []
    for i in range(num_stages):
        if i > 0:
            points, lengths, normals = grid_subsample(points, lengths, normals, voxel_size=voxel_size)
        if i == num_stages - 1:
            # data.py:34-43: at most 2000 superpoints per cloud (pair mode only, as in the reference)
            if lengths.shape[0] == 2:
                l0, l1 = (int(v) for v in lengths.tolist())
                if l0 > 2000 or l1 > 2000:
                    k0, k1 = min(l0, 2000), min(l1, 2000)
                    points = torch.cat((points[:k0], points[l0:l0 + k1]), dim=0)
                    normals = torch.cat((normals[:k0], normals[l0:l0 + k1]), dim=0)
                    lengths = torch.tensor([k0, k1], dtype=torch.int64, device=points.device)
            elif ext.LAST.get('max_length', 1 << 30) > 2000 and bool((lengths > 2000).any()):
                # stacked pairs: the same cap per cloud, so that a pair gives the same result batched and alone (the
                # size of the largest cloud came back with grid_subsample's own status read-back: no extra sync unless a
                # cloud really is over the cap; the slicing itself is the rare path)
                host = [int(v) for v in lengths.tolist()]
                keep, start = [], 0
                for l in host:
                    keep.append(torch.arange(start, start + min(l, 2000), device=points.device))
                    start += l
                keep = torch.cat(keep)
                points, normals = points[keep].contiguous(), normals[keep].contiguous()
                lengths = torch.tensor([min(l, 2000) for l in host], dtype=torch.int64, device=points.device)
        points_list.append(points)
        lengths_list.append(lengths)
        normals_list.append(normals)
        voxel_size *= 2

    neighbors_list, subsampling_list, upsampling_list = [], [], []
    subsampling_width = []
    pending = []  # (list, index, limit, status)

    def search(dst, q, s, ql, sl, r, limit):
        if limit <= 0:
            from .ops import radius_search
            dst.append(radius_search(q, s, ql, sl, r, limit))
            return
        out, status = radius_search_deferred(q, s, ql, sl, r, limit)
        dst.append(out)
        pending.append((dst, len(dst) - 1, limit, status))
        if batched and dst is subsampling_list:
            cloud_max = status[_lib.SE3ET_STATUS_WORDS:]
            subsampling_width.append(cloud_max.view(num_pairs, 2).amax(dim=1).clamp(max=limit).contiguous())

    for i in range(num_stages):
        cur_points, cur_lengths = points_list[i], lengths_list[i]
        search(neighbors_list, cur_points, cur_points, cur_lengths, cur_lengths, radius, neighbor_limits[i])
        if i < num_stages - 1:
            sub_points, sub_lengths = points_list[i + 1], lengths_list[i + 1]
            search(subsampling_list, sub_points, cur_points, sub_lengths, cur_lengths, radius, neighbor_limits[i])
            if backbone_only and i == 0:
                upsampling_list.append(None)
            else:
                search(upsampling_list, cur_points, sub_points, cur_lengths, sub_lengths, radius * 2,
                       1 if backbone_only else neighbor_limits[i + 1])
        radius *= 2

    if pending and not batched:
        # one transfer for all searches: width is min(max_count, limit) as in ops/radius_search.py:24-27
        stats = torch.stack([p[3] for p in pending]).cpu()
        for (dst, idx, limit, _), st in zip(pending, stats):
            max_count = int(st[_lib.STATUS_MAX_COUNT])
            if max_count < limit:
                dst[idx] = dst[idx][:, :max_count].contiguous()

    extra = {}
    if batched:
        zero = lengths_list[0].new_zeros((1,))
        extra['pair_offsets'] = [torch.cat([zero, l.view(num_pairs, 2).sum(dim=1).cumsum(0)]) for l in lengths_list]
        extra['subsampling_width'] = subsampling_width
    return {
        **extra,
        'points': points_list,
        'lengths': lengths_list,
        'neighbors': neighbors_list,
        'subsampling': subsampling_list,
        'upsampling': upsampling_list,
        'normals': normals_list,
    }
