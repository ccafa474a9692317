"""Generates tests/golden/lgr_ref.npz from the UNMODIFIED reference modules
geotransformer.modules.geotransformer.LocalGlobalRegistration and geotransformer.modules.registration.weighted_procrustes
(CPU; `.cuda()` neutralised by the shim).

    python tests/golden/make_lgr_golden.py        (needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_import_shim as shim  # noqa: E402

shim.install("se3eti.3dmatch")
import geotransformer.modules.geotransformer  # noqa: F401,E402
from geotransformer.modules.geotransformer import LocalGlobalRegistration  # noqa: E402
from geotransformer.modules.registration import weighted_procrustes  # noqa: E402


def rot(rng):
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_case(rng, B, K, noise, outlier_frac, empty_patches=0):
    """B patch pairs of K points: src = R^T (ref - t) + noise for matching slots, log scores peaked on the true
    permutation, a fraction of patches scrambled (outliers), masks with padding."""
    R, t = rot(rng), rng.normal(size=3) * 0.5
    ref = rng.uniform(-1.5, 1.5, size=(B, K, 3)).astype(np.float32)
    perm = np.stack([rng.permutation(K) for _ in range(B)])
    src = np.zeros_like(ref)
    for b in range(B):
        p = (ref[b] - t) @ R                      # R^T (ref - t)
        if rng.random() < outlier_frac:
            p = rng.uniform(-1.5, 1.5, size=(K, 3))
        src[b, perm[b]] = p + rng.normal(size=(K, 3)) * noise
    ref_masks = rng.random((B, K)) > 0.15
    src_masks = rng.random((B, K)) > 0.15
    logits = rng.normal(size=(B, K, K)).astype(np.float32) * 1.5 - 6.0
    for b in range(B):
        good = rng.random(K) < 0.7
        logits[b, np.arange(K)[good], perm[b][good]] = rng.uniform(-1.2, -0.05, size=int(good.sum()))
    for b in range(empty_patches):
        logits[b] = -20.0
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, t
    return ref, src.astype(np.float32), ref_masks, src_masks, logits, T.astype(np.float32)


def main():
    rng = np.random.default_rng(11)
    out = {}
    # weighted_procrustes: single and batched, with zero / sub-threshold weights
    for tag, n in (("wp_small", 7), ("wp_large", 300)):
        R, t = rot(rng), rng.normal(size=3)
        src = rng.normal(size=(n, 3)).astype(np.float32)
        ref = (src @ R.T + t + rng.normal(size=(n, 3)) * 0.01).astype(np.float32)
        w = rng.random(n).astype(np.float32)
        w[::5] = 0.0
        T = weighted_procrustes(torch.from_numpy(src), torch.from_numpy(ref), torch.from_numpy(w), return_transform=True)
        out[tag + "_src"], out[tag + "_ref"], out[tag + "_w"], out[tag + "_T"] = src, ref, w, T.numpy()
    cases = {"clean": (24, 16, 0.005, 0.2, 0), "noisy": (40, 64, 0.02, 0.4, 3), "degenerate": (6, 8, 0.01, 0.0, 6)}
    for tag, (B, K, noise, outl, empty) in cases.items():
        ref, src, rm, sm, logits, T_gt = make_case(rng, B, K, noise, outl, empty)
        if tag == "degenerate":
            # at most two correspondences per patch: no chunk reaches correspondence_threshold
            logits[:] = -20.0
            for b in range(B):
                logits[b, b % K, (b + 1) % K] = -0.1
                rm[b, b % K] = sm[b, (b + 1) % K] = True
        lgr = LocalGlobalRegistration(3, 0.1, mutual=True, confidence_threshold=0.05, use_dustbin=False,
                                      use_global_score=False, correspondence_threshold=3, correspondence_limit=None,
                                      num_refinement_steps=5)
        with torch.no_grad():
            rp, sp, cs, T = lgr(torch.from_numpy(ref), torch.from_numpy(src), torch.from_numpy(rm), torch.from_numpy(sm),
                                torch.from_numpy(logits), torch.ones(B))
        for k, v in (("ref", ref), ("src", src), ("rm", rm), ("sm", sm), ("logits", logits), ("T_gt", T_gt),
                     ("out_ref", rp.numpy()), ("out_src", sp.numpy()), ("out_scores", cs.numpy()), ("out_T", T.numpy())):
            out["%s_%s" % (tag, k)] = v
        print(tag, "correspondences", len(cs), "transform error", np.abs(T.numpy() - T_gt).max())
    np.savez_compressed(os.path.join(HERE, "lgr_ref.npz"), **out)


if __name__ == "__main__":
    main()
