"""oracle/calibrate.py against the unmodified reference `calibrate_neighbors_stack_mode` (tests/golden/calibrate_ref.npz,
made by tests/golden/make_calibrate_golden.py from the same synthetic pairs)."""
import os

import numpy as np
import pytest

from oracle import calibrate as ocal
from se3et_b200 import synthetic

CASES = {
    "tdm_small": (lambda: [synthetic.make_3dmatch_pair(s, crop=0.9) for s in (3, 5, 13)]),
    "tdm_loose": (lambda: [synthetic.make_3dmatch_pair(s, crop=0.9) for s in (13, 5)]),
    "kitti_small": (lambda: [synthetic.make_kitti_pair(s, target_points=4000) for s in (0, 1)]),
}


def clouds_of(name):
    return [(p["ref_points"].astype(np.float32), p["src_points"].astype(np.float32)) for p in CASES[name]()]


@pytest.mark.parametrize("name", sorted(CASES))
def test_calibration_matches_reference(golden_dir, name):
    gold = np.load(os.path.join(golden_dir, "calibrate_ref.npz"))
    stages, voxel, radius, keep, thresh = gold[name + "_params"]
    clouds = clouds_of(name)
    assert np.array_equal(gold[name + "_seeds"], [len(c[0]) for c in clouds] + [len(c[1]) for c in clouds])
    got = ocal.calibrate_neighbors_stack_mode(clouds, int(stages), float(voxel), float(radius), float(keep), int(thresh))
    assert np.array_equal(got, gold[name + "_limits"]), (got, gold[name + "_limits"])
