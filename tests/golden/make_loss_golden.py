"""Generates tests/golden/loss_ref.npz from the UNMODIFIED reference: weighted_circle_loss
(geotransformer/modules/loss/circle_loss.py), get_node_correspondences (modules/registration/matching.py),
CoarseMatchingLoss / FineMatchingLoss (experiments/se3eti.3dmatch/loss.py) and LearnableLogOptimalTransport gradients.

    python tests/golden/make_loss_golden.py        (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import_shim as shim  # noqa: E402

cfg = shim.make_cfg("se3eti.3dmatch")
import geotransformer.modules.geotransformer  # noqa: F401,E402
from geotransformer.modules.registration.matching import get_node_correspondences  # noqa: E402
from geotransformer.modules.sinkhorn import LearnableLogOptimalTransport  # noqa: E402
import loss as ref_loss  # noqa: E402  (experiments/se3eti.3dmatch/loss.py)


def main():
    g = torch.Generator().manual_seed(3)
    out = {}
    # patches: M reference / N source nodes with K points each around the node, a planted rigid motion
    M, N, K = 14, 11, 16
    q = torch.randn(4, generator=g)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    R = torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                      [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                      [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    T = torch.eye(4)
    T[:3, :3], T[:3, 3] = R, torch.tensor([0.3, -0.2, 0.1])
    ref_nodes = torch.rand(M, 3, generator=g) * 0.6
    ref_knn = ref_nodes[:, None, :] + torch.randn(M, K, 3, generator=g) * 0.04
    # source nodes: the first N reference nodes moved back by T^-1, points jittered
    src_nodes = (ref_nodes[:N] - T[:3, 3]) @ R
    src_knn = (ref_knn[:N] - T[:3, 3]) @ R + torch.randn(N, K, 3, generator=g) * 0.01
    ref_masks, src_masks = torch.rand(M, generator=g) > 0.1, torch.rand(N, generator=g) > 0.1
    ref_km, src_km = torch.rand(M, K, generator=g) > 0.2, torch.rand(N, K, generator=g) > 0.2
    ref_km[:, 0] = True
    src_km[:, 0] = True
    idx, ov = get_node_correspondences(ref_nodes, src_nodes, ref_knn, src_knn, T, 0.05, ref_masks, src_masks, ref_km, src_km)
    for k, v in dict(ref_nodes=ref_nodes, src_nodes=src_nodes, ref_knn=ref_knn, src_knn=src_knn, T=T, ref_masks=ref_masks,
                     src_masks=src_masks, ref_km=ref_km, src_km=src_km, gt_idx=idx, gt_ov=ov).items():
        out["nc_" + k] = v.numpy()
    print("node correspondences", idx.shape[0])
    # coarse loss + its gradient
    ref_f = torch.nn.functional.normalize(torch.randn(M, 32, generator=g), dim=1).requires_grad_(True)
    src_f = torch.nn.functional.normalize(torch.randn(N, 32, generator=g), dim=1).requires_grad_(True)
    cl = ref_loss.CoarseMatchingLoss(cfg)
    l = cl({"ref_feats_c": ref_f, "src_feats_c": src_f, "gt_node_corr_indices": idx, "gt_node_corr_overlaps": ov})
    l.backward()
    out.update(cl_ref=ref_f.detach().numpy(), cl_src=src_f.detach().numpy(), cl_loss=l.detach().numpy(),
               cl_gref=ref_f.grad.numpy(), cl_gsrc=src_f.grad.numpy())
    # optimal transport + fine loss + gradients (scores and alpha)
    B = 5
    scores = (torch.randn(B, K, K, generator=g) * 1.5).requires_grad_(True)
    ot = LearnableLogOptimalTransport(20)
    with torch.no_grad():
        ot.alpha.fill_(0.4)
    rm, cm = ref_km[:B], src_km[:B]
    ms = ot(scores, rm, cm)
    fl = ref_loss.FineMatchingLoss(cfg)
    l2 = fl({"ref_node_corr_knn_points": ref_knn[:B], "src_node_corr_knn_points": src_knn[:B],
             "ref_node_corr_knn_masks": rm, "src_node_corr_knn_masks": cm, "matching_scores": ms}, {"transform": T})
    l2.backward()
    out.update(fl_scores=scores.detach().numpy(), fl_ms=ms.detach().numpy(), fl_loss=l2.detach().numpy(),
               fl_gscores=scores.grad.numpy(), fl_galpha=ot.alpha.grad.numpy())
    print("coarse loss", float(l), "fine loss", float(l2))
    np.savez_compressed(os.path.join(HERE, "loss_ref.npz"), **out)


if __name__ == "__main__":
    main()
