"""Generates tests/golden/calibrate_ref.npz from the UNMODIFIED reference function
geotransformer.utils.data.calibrate_neighbors_stack_mode driven by the reference's own collate function
(registration_collate_fn_stack_mode) over a small list dataset of synthetic pairs; geotransformer.ext comes from the
reference C++ built in oracle/_ref (ref_import_shim.py).

    python tests/golden/make_calibrate_golden.py        (needs /root/reference)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import_shim as shim  # noqa: E402

shim.install("se3eti.3dmatch")
from geotransformer.utils.data import calibrate_neighbors_stack_mode, registration_collate_fn_stack_mode  # noqa: E402
from se3et_b200 import synthetic  # noqa: E402


def dataset(pairs):
    items = []
    for ref, src in pairs:
        items.append({"ref_points": ref, "src_points": src, "ref_feats": np.ones((len(ref), 1), np.float32),
                      "src_feats": np.ones((len(src), 1), np.float32), "transform": np.eye(4, dtype=np.float32)})
    return items


def main():
    out = {}
    cases = {
        # name: (pairs, num_stages, voxel, radius, keep_ratio, sample_threshold)
        "tdm_small": ([synthetic.make_3dmatch_pair(s, crop=0.9) for s in (3, 5, 13)], 4, 0.025, 0.0625, 0.8, 2000),
        "tdm_loose": ([synthetic.make_3dmatch_pair(s, crop=0.9) for s in (13, 5)], 3, 0.025, 0.0625, 0.95, 10 ** 9),
        "kitti_small": ([synthetic.make_kitti_pair(s, target_points=4000) for s in (0, 1)], 5, 0.3, 4.25 * 0.3, 0.8, 2000),
    }
    for name, (pairs, stages, voxel, radius, keep, thresh) in cases.items():
        clouds = [(p["ref_points"].astype(np.float32), p["src_points"].astype(np.float32)) for p in pairs]
        limits = calibrate_neighbors_stack_mode(dataset(clouds), registration_collate_fn_stack_mode, stages, voxel, radius,
                                                keep_ratio=keep, sample_threshold=thresh)
        print(name, limits)
        out[name + "_limits"] = np.asarray(limits, dtype=np.int64)
        out[name + "_params"] = np.array([stages, voxel, radius, keep, thresh], dtype=np.float64)
        out[name + "_seeds"] = np.array([len(c[0]) for c in clouds] + [len(c[1]) for c in clouds], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, "calibrate_ref.npz"), **out)


if __name__ == "__main__":
    main()
