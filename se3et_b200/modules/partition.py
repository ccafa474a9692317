"""Drop-in for geotransformer/modules/ops/pointcloud_partition.py:60-107 on the CUDA path."""
import torch

from ..ops.partition_ops import point_to_node_partition_stacked


@torch.no_grad()
def point_to_node_partition(points, nodes, point_limit, return_count=False):
    """Same signature and return order as the reference: point_to_node (N,), [node_sizes (M,),] node_masks (M,),
    node_knn_indices (M, K), node_knn_masks (M, K).  Ties (equal fp32 distances) are ordered by index; the reference leaves
    them to torch.min / torch.topk."""
    dev = points.device
    pl = torch.tensor([points.shape[0]], dtype=torch.int64, device=dev)
    nl = torch.tensor([nodes.shape[0]], dtype=torch.int64, device=dev)
    res = point_to_node_partition_stacked(points.float(), pl, nodes.float(), nl, point_limit, return_count=return_count)
    if return_count:
        p2n, masks, knn, knn_masks, sizes = res
        return p2n, sizes, masks, knn, knn_masks
    return res
