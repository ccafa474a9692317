"""GPU parity of the fine-stage registration kernels (csrc/registration.cu) through the C ABI: against the numpy oracle
(oracle/registration.py) and the outputs of the unmodified reference modules (tests/golden/lgr_ref.npz).
Correspondence sets are exact; transforms agree to 1e-4 (fp64 Jacobi vs the reference's fp32 sums + LAPACK SVD)."""
import os

import numpy as np
import pytest
import torch

from oracle import registration as oreg
from se3et_b200.modules.registration import LocalGlobalRegistration, WeightedProcrustes, weighted_procrustes

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "lgr_ref.npz"))


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.mark.parametrize("tag", ["wp_small", "wp_large"])
def test_weighted_procrustes_matches_reference(gold, tag):
    T = weighted_procrustes(t(gold[tag + "_src"]), t(gold[tag + "_ref"]), t(gold[tag + "_w"]), return_transform=True)
    assert np.abs(T.cpu().numpy() - gold[tag + "_T"]).max() < 5e-5
    R, tr = WeightedProcrustes()(t(gold[tag + "_src"])[None].repeat(3, 1, 1), t(gold[tag + "_ref"])[None].repeat(3, 1, 1),
                                 t(gold[tag + "_w"])[None].repeat(3, 1))
    assert R.shape == (3, 3, 3) and np.abs(R[2].cpu().numpy() - gold[tag + "_T"][:3, :3]).max() < 5e-5
    assert np.abs(tr[1].cpu().numpy() - gold[tag + "_T"][:3, 3]).max() < 5e-5


def test_weighted_procrustes_reflection_and_degenerate_inputs():
    rng = np.random.default_rng(5)
    src = rng.normal(size=(50, 3)).astype(np.float32)
    mirror = src * np.array([1, 1, -1], np.float32)          # best orthogonal map is a reflection: det must stay +1
    for ref in (mirror, np.zeros_like(src), src[:, [1, 2, 0]] + 3.0):
        T = weighted_procrustes(t(src), t(ref), None, return_transform=True).cpu().numpy().astype(np.float64)
        want = oreg.weighted_procrustes(src, ref)
        R = T[:3, :3]
        assert abs(np.linalg.det(R) - 1) < 1e-5 and np.allclose(R @ R.T, np.eye(3), atol=1e-5)
        # the optimum is unique unless singular values tie: compare the objective, not the matrix
        def cost(M):
            return np.sum((src @ M[:3, :3].T + M[:3, 3] - ref) ** 2)
        assert cost(T) <= cost(want) * (1 + 1e-4) + 1e-6


@pytest.mark.parametrize("tag", ["clean", "noisy", "degenerate"])
def test_local_global_registration_matches_reference_and_oracle(gold, tag):
    lgr = LocalGlobalRegistration(3, 0.1, mutual=True, confidence_threshold=0.05, correspondence_threshold=3,
                                  num_refinement_steps=5)
    args = [gold[tag + "_" + k] for k in ("ref", "src", "rm", "sm", "logits")]
    rp, sp, sc, T = lgr(*[t(a) for a in args], None)
    w_rp, w_sp, w_sc, w_T = oreg.local_global_registration(*args)
    assert np.array_equal(rp.cpu().numpy(), w_rp) and np.array_equal(sp.cpu().numpy(), w_sp)
    assert np.array_equal(rp.cpu().numpy(), gold[tag + "_out_ref"])
    assert np.allclose(sc.cpu().numpy(), w_sc, rtol=1e-5, atol=0)
    assert np.abs(T.cpu().numpy() - w_T).max() < 1e-4
    assert np.abs(T.cpu().numpy() - gold[tag + "_out_T"]).max() < 1e-4


def test_stacked_pairs_register_independently(gold):
    """Several pairs in one call (patch_offsets), with a trailing dustbin row / column in the score matrices and an empty
    pair in the middle: every pair's output equals its own single-pair run."""
    lgr = LocalGlobalRegistration(3, 0.1)
    tags = ["noisy", "clean", "degenerate"]
    K = 64
    parts, offs = [], [0]
    for tag in tags:
        ref, src, rm, sm, lg = [gold[tag + "_" + k] for k in ("ref", "src", "rm", "sm", "logits")]
        b, k = rm.shape
        pad = lambda a, fill: np.concatenate([a, np.full((b, K - k) + a.shape[2:], fill, a.dtype)], 1)
        lg2 = np.full((b, K + 1, K + 1), -30.0, np.float32)
        lg2[:, :k, :k] = lg
        parts.append((pad(ref, 0), pad(src, 0), pad(rm, False), pad(sm, False), lg2))
        offs.append(offs[-1] + b)
        if tag == "clean":
            offs.append(offs[-1])  # an empty pair
    cat = [np.concatenate([p[i] for p in parts]) for i in range(5)]
    rp, sp, sc, coff, T = lgr.forward_pairs(*[t(a) for a in cat], torch.tensor(offs, device=DEV))
    coff = coff.cpu().numpy()
    assert len(coff) == 5 and coff[2] == coff[3]
    assert np.allclose(T[2].cpu().numpy(), np.eye(4), atol=1e-6)   # no correspondences: identity
    for slot, tag in ((0, "noisy"), (1, "clean"), (3, "degenerate")):
        assert np.array_equal(rp[coff[slot]:coff[slot + 1]].cpu().numpy(), gold[tag + "_out_ref"])
        assert np.abs(T[slot].cpu().numpy() - gold[tag + "_out_T"]).max() < 1e-4
