"""Generates tests/golden/partition_ref.npz from the UNMODIFIED reference function
geotransformer.modules.ops.pointcloud_partition.point_to_node_partition (run on the CPU: `.cuda()` is neutralised by
ref_import_shim).  Inputs: the fine (stage 2) and coarse (last stage) levels of the committed point pyramids.

    python tests/golden/make_partition_golden.py        (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import_shim as shim  # noqa: E402  (must precede any geotransformer import)

shim.install("se3eti.3dmatch")
import geotransformer.modules.geotransformer  # noqa: F401,E402  (import order: avoids the reference's circular import)
from geotransformer.modules.ops.pointcloud_partition import point_to_node_partition  # noqa: E402


def main():
    out = {}
    for tag, fname, limit in (("demo", "points_demo_crop.npz", 64), ("demo8", "points_demo_crop.npz", 8),
                              ("synth", "points_synth_small.npz", 16)):
        g = np.load(os.path.join(HERE, fname))
        stages = len([k for k in g.files if k.startswith("points_")])
        pf, lf = g["points_1"], g["lengths_1"]
        pc, lc = g["points_%d" % (stages - 1)], g["lengths_%d" % (stages - 1)]
        po = no = 0
        for b in range(len(lf)):
            pts = torch.from_numpy(pf[po:po + lf[b]].astype(np.float32))
            nodes = torch.from_numpy(pc[no:no + lc[b]].astype(np.float32))
            p2n, sizes, masks, knn, knn_masks = point_to_node_partition(pts, nodes, limit, return_count=True)
            key = "%s_%d_" % (tag, b)
            out[key + "points"], out[key + "nodes"] = pts.numpy(), nodes.numpy()
            out[key + "limit"] = np.int64(limit)
            out[key + "point_to_node"] = p2n.numpy().astype(np.int64)
            out[key + "node_sizes"] = sizes.numpy().astype(np.int64)
            out[key + "node_masks"] = masks.numpy()
            # the reference leaves masked slots' indices at N and does not order ties: store the canonical form
            k_idx, k_msk = knn.numpy().astype(np.int64), knn_masks.numpy()
            out[key + "node_knn_indices"] = k_idx
            out[key + "node_knn_masks"] = k_msk
            po += lf[b]
            no += lc[b]
            print(key, pts.shape, nodes.shape, "max node size", int(sizes.max()))
    np.savez_compressed(os.path.join(HERE, "partition_ref.npz"), **out)


if __name__ == "__main__":
    main()
