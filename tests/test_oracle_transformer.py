"""CPU: pins oracle/transformer.py (geometric embedding, SE3ET-I transformer, SuperPointMatching) against fixtures
from the unmodified reference modules (tests/golden/make_model_golden.py)."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import points as op
from oracle import transformer as ot


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "model_small.npz"))


def transformer_state_dict(in_dim=256):
    S = helpers.SMALL_CFG
    c = S["hidden_dim"]
    shapes = {}

    def lin(name, cin, cout):
        shapes[name + ".weight"] = (cout, cin)
        shapes[name + ".bias"] = (cout,)

    def norm(name):
        shapes[name + ".weight"] = (c,)
        shapes[name + ".bias"] = (c,)

    lin("transformer.embedding.proj_d", c, c)
    lin("transformer.embedding.proj_a", c, c)
    lin("transformer.in_proj", in_dim, c)
    lin("transformer.out_proj", c, S["tr_output_dim"])
    for i, block in enumerate(S["blocks"]):
        base = "transformer.transformer.layers.%d" % i
        for pn in ("proj_q", "proj_k", "proj_v") + (("proj_p",) if "self" in block else ()):
            lin(base + ".attention.attention." + pn, c, c)
        lin(base + ".attention.linear", c, c)
        norm(base + ".attention.norm")
        lin(base + ".output.expand", c, 2 * c)
        lin(base + ".output.squeeze", 2 * c, c)
        norm(base + ".output.norm")
    return {k: helpers.seeded_tensor(k, s) for k, s in shapes.items()}


def coarse_inputs(gold):
    S = helpers.SMALL_CFG
    d = op.precompute_data_stack_mode(gold["in_points"], gold["in_lengths"], 4, S["init_voxel"], S["init_radius"],
                                      [38, 36, 36, 38], impl="oracle")
    n = int(d["lengths"][3][0])
    pc = torch.from_numpy(d["points"][3])
    fc = torch.from_numpy(gold["feats_c"])
    return pc[:n], pc[n:], fc[:n], fc[n:]


def test_transformer_matches_reference(gold):
    S = helpers.SMALL_CFG
    rp, sp, rf, sf = coarse_inputs(gold)
    sd = transformer_state_dict()
    r, s, e0, _ = ot.geometric_transformer(sd, rp, sp, rf, sf, S["blocks"], S["hidden_dim"], S["num_heads"],
                                           S["sigma_d"], S["sigma_a"], S["angle_k"])
    assert torch.allclose(e0, torch.from_numpy(gold["ref_embedding"].astype(np.float32)), rtol=2e-3, atol=2e-3)
    assert torch.allclose(r, torch.from_numpy(gold["ref_feats_c"]), rtol=1e-3, atol=1e-4), \
        (r - torch.from_numpy(gold["ref_feats_c"])).abs().max()
    assert torch.allclose(s, torch.from_numpy(gold["src_feats_c"]), rtol=1e-3, atol=1e-4)


def test_superpoint_matching_matches_reference(gold):
    ri, si, sc = ot.superpoint_matching(torch.from_numpy(gold["spm_ref_feats"]), torch.from_numpy(gold["spm_src_feats"]),
                                        torch.from_numpy(gold["spm_ref_masks"]), torch.from_numpy(gold["spm_src_masks"]),
                                        64, True)
    assert torch.allclose(sc, torch.from_numpy(gold["spm_scores"]), rtol=1e-5, atol=1e-9)
    # torch.topk's tie order is unspecified: compare as sets of (ref, src) pairs, and exactly where scores are distinct
    want = set(zip(gold["spm_ref_idx"].tolist(), gold["spm_src_idx"].tolist()))
    got = set(zip(ri.tolist(), si.tolist()))
    assert got == want
    distinct = np.concatenate([[True], np.abs(np.diff(gold["spm_scores"])) > 1e-7 * gold["spm_scores"][:-1]])
    distinct &= np.concatenate([distinct[1:], [True]])
    assert np.array_equal(ri.numpy()[distinct], gold["spm_ref_idx"][distinct])
    assert np.array_equal(si.numpy()[distinct], gold["spm_src_idx"][distinct])
    assert not (set(ri.tolist()) & {3}) and not (set(si.tolist()) & {0, 7})  # masked superpoints never appear


# ---- SE3ET-E block list (cross_a_soft / cross_r_soft / invariant self + cross), fixture from the reference ----------
BLOCKS_E = ['self_eq', 'cross_a_soft', 'self_eq', 'cross_r_soft', 'self', 'cross']


@pytest.fixture(scope="module")
def gold_e(golden_dir):
    return np.load(os.path.join(golden_dir, "model_e_small.npz"))


def transformer_e_state_dict(gold_e, prefix="transformer."):
    """Seeded parameters under the reference's own key names (shapes recorded in the fixture)."""
    sd = {}
    for item in gold_e["param_shapes"]:
        name, shape = str(item).split(":")
        if helpers.is_constant(name) or "anchors" in name:
            continue
        shape = tuple(int(v) for v in shape.split(",")) if shape else ()
        sd[prefix + name] = helpers.seeded_tensor("transformer_e." + name, shape)
    return sd


def test_rotation_permutations_match_reference(gold_e):
    ours = {tuple(r) for r in ot.octahedral_rotation_perms().tolist()}
    ref = {tuple(r) for r in gold_e["trace_idx_ori"].tolist()}
    assert ours == ref and len(ours) == 24
    from se3et_b200.modules import octahedral
    assert np.allclose(octahedral.tables()["anchors"], np.transpose(gold_e["anchors_embedding"], (0, 2, 1)), atol=1e-6)


@pytest.mark.parametrize("tag,nlev", [("nosh", 0), ("sh", 2)])
def test_transformer_e_matches_reference(gold, gold_e, tag, nlev):
    S = helpers.SMALL_CFG
    rp, sp, rf, sf = coarse_inputs(gold)
    sd = transformer_e_state_dict(gold_e)
    if nlev == 0:
        sd = {k: v for k, v in sd.items() if "proj_eq" not in k}
    anchors = torch.from_numpy(np.transpose(gold_e["anchors_embedding"], (0, 2, 1)).copy()).float()
    r, s = ot.geometric_transformer_eq(sd, rp, sp, rf, sf, BLOCKS_E, S["hidden_dim"], S["num_heads"], S["sigma_d"],
                                       S["sigma_a"], S["angle_k"], anchors, n_level_equiv=nlev)
    assert torch.allclose(r, torch.from_numpy(gold_e["ref_feats_" + tag]), rtol=1e-3, atol=2e-4)
    assert torch.allclose(s, torch.from_numpy(gold_e["src_feats_" + tag]), rtol=1e-3, atol=2e-4)
