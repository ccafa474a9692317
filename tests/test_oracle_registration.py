"""The numpy restatement of the fine-stage registration (oracle/registration.py) against the outputs of the unmodified
reference modules (tests/golden/lgr_ref.npz, made by tests/golden/make_lgr_golden.py)."""
import os

import numpy as np
import pytest

from oracle import registration as oreg


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "lgr_ref.npz"))


@pytest.mark.parametrize("tag", ["wp_small", "wp_large"])
def test_weighted_procrustes_matches_reference(gold, tag):
    T = oreg.weighted_procrustes(gold[tag + "_src"], gold[tag + "_ref"], gold[tag + "_w"])
    assert np.abs(T - gold[tag + "_T"]).max() < 2e-5
    R = T[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-10) and np.linalg.det(R) > 0.999


@pytest.mark.parametrize("tag", ["clean", "noisy", "degenerate"])
def test_local_global_registration_matches_reference(gold, tag):
    rp, sp, sc, T = oreg.local_global_registration(gold[tag + "_ref"], gold[tag + "_src"], gold[tag + "_rm"],
                                                   gold[tag + "_sm"], gold[tag + "_logits"])
    # the correspondence set is exact (same order: patch, reference slot, source slot)
    assert np.array_equal(rp, gold[tag + "_out_ref"]) and np.array_equal(sp, gold[tag + "_out_src"])
    assert np.allclose(sc, gold[tag + "_out_scores"], rtol=1e-6, atol=0)
    assert np.abs(T - gold[tag + "_out_T"]).max() < 5e-5
    if tag != "degenerate":
        assert np.abs(T - gold[tag + "_T_gt"]).max() < 5e-3   # and both recover the planted motion
