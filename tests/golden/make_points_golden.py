"""Generates tests/golden/points_*.npz from the UNMODIFIED reference C++ (oracle/_ref, built from
/root/reference by oracle/Makefile), canonicalised per SURVEY.md 8(c).  Run in the build container:

    python tests/golden/make_points_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import points as op  # noqa: E402
from se3et_b200 import synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pyramid(points, lengths, stages, voxel, radius, limits):
    d = op.precompute_data_stack_mode(points, lengths, stages, voxel, radius, limits, impl="ref")
    out = {"in_points": points.astype(np.float32), "in_lengths": lengths.astype(np.int64),
           "voxel": np.float64(voxel), "radius": np.float64(radius), "limits": np.asarray(limits, np.int64)}
    for i in range(stages):
        out["points_%d" % i] = d["points"][i]
        out["lengths_%d" % i] = d["lengths"][i]
        out["neighbors_%d" % i] = d["neighbors"][i].astype(np.int32)
        if i < stages - 1:
            out["subsampling_%d" % i] = d["subsampling"][i].astype(np.int32)
            out["upsampling_%d" % i] = d["upsampling"][i].astype(np.int32)
    return out


def main():
    assert op.have_ref() or os.path.isdir("/root/reference"), "needs the reference build"
    # (1) a crop of the reference's own demo pair (data/demo/{ref,src}.npy)
    ref = np.load("/root/reference/data/demo/ref.npy").astype(np.float32)
    src = np.load("/root/reference/data/demo/src.npy").astype(np.float32)
    ref = ref[(ref[:, 0] < -0.3) & (ref[:, 1] < -0.5)]
    src = src[(src[:, 0] < -0.5) & (src[:, 1] < -0.3)]
    pts = np.concatenate([ref, src])
    lens = np.array([len(ref), len(src)])
    np.savez_compressed(os.path.join(HERE, "points_demo_crop.npz"),
                        **pyramid(pts, lens, 4, 0.025, 0.0625, [38, 36, 36, 38]))
    # (2) a small synthetic 3DMatch-shaped pair, unlimited neighbour width at every stage
    p = synthetic.make_3dmatch_pair(7, target_points=1200)
    pts = np.concatenate([p["ref_points"], p["src_points"]])
    lens = np.array([len(p["ref_points"]), len(p["src_points"])])
    np.savez_compressed(os.path.join(HERE, "points_synth_small.npz"), **pyramid(pts, lens, 3, 0.025, 0.0625, [0, 0, 0]))
    for f in ("points_demo_crop.npz", "points_synth_small.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
