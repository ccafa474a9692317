"""CPU: pins the torch restatement of the E2PN backbone (oracle/e2pn.py) against fixtures produced by the
unmodified reference modules (tests/golden/make_model_golden.py)."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import e2pn as oe
from oracle import points as op


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "model_small.npz"))


@pytest.fixture(scope="module")
def pyramid(gold):
    S = helpers.SMALL_CFG
    return op.precompute_data_stack_mode(gold["in_points"], gold["in_lengths"], 4, S["init_voxel"], S["init_radius"],
                                         [38, 36, 36, 38], impl="oracle")


def test_group_tables_match_reference(gold):
    t = oe.octahedral_tables()
    assert t["k_real"] == 6
    assert np.allclose(t["anchors"], gold["const_anchors"], atol=1e-6)
    assert np.allclose(t["quotient"], gold["const_quotient_anchors"], atol=1e-6)
    radius = helpers.SMALL_CFG["init_radius"]
    assert np.allclose(t["kp_unit"] * 0.7 * radius, gold["const_kernel_points"], atol=1e-7)
    # reference buffers are (K, A, R) expansions: kidx_rot[k, :, r], ridx_rot[:, a, r]
    assert np.array_equal(t["kidx"], gold["const_kidx_rot"][:, 0, :])
    assert np.array_equal(t["ridx"], gold["const_ridx_rot"][0])
    assert (gold["const_kidx_rot"] == gold["const_kidx_rot"][:, :1, :]).all()
    assert (gold["const_ridx_rot"] == gold["const_ridx_rot"][:1]).all()


def _conv_params(gold):
    names = {"conv.weights": (6, 6, 8, 16)}
    return {k: helpers.seeded_tensor(k, s) for k, s in names.items()}


def test_kpconv_matches_reference(gold, pyramid):
    t = oe.octahedral_tables()
    p1 = torch.from_numpy(pyramid["points"][1])
    nb1 = torch.from_numpy(pyramid["neighbors"][1])
    x = helpers.seeded_tensor("conv.input", (p1.shape[0], 6, 8))
    w = _conv_params(gold)["conv.weights"]
    kp = torch.from_numpy(gold["const_kernel_points"])
    out = oe.kpconv_inter_so3(p1, p1, nb1, x, w, kp, helpers.SMALL_CFG["init_sigma"], t["kidx"], t["ridx"])
    assert torch.allclose(out, torch.from_numpy(gold["conv_out"]), rtol=1e-4, atol=1e-5)


def backbone_state_dict():
    """Shapes of the reduced-width E2PN (init_dim 16, output_dim 32), values from the seeded scheme."""
    S = helpers.SMALL_CFG
    d, g = S["init_dim"], S["group_norm"]
    shapes = {}

    def norm(prefix, c):
        shapes[prefix + ".norm.weight"] = (c,)
        shapes[prefix + ".norm.bias"] = (c,)

    def unary(prefix, cin, cout):
        shapes[prefix + ".mlp.weight"] = (cout, cin)
        shapes[prefix + ".mlp.bias"] = (cout,)
        norm(prefix + ".norm", cout)

    def conv(prefix, cin, cout):
        shapes[prefix + ".interso3.conv.weights"] = (6, 6, cin, cout)
        norm(prefix + ".interso3.norm", cout)
        norm(prefix + ".norm", cout)

    def resnet(prefix, cin, cout):
        mid = cout // 4
        if cin != mid:
            unary(prefix + ".unary1", cin, mid)
        conv(prefix, mid, mid)
        unary(prefix + ".unary2", mid, cout)
        if cin != cout:
            unary(prefix + ".skip_conv", cin, cout)

    conv("backbone.encoder1_1", 1, d)
    resnet("backbone.encoder1_2", d, 2 * d)
    resnet("backbone.encoder2_1", 2 * d, 2 * d)
    resnet("backbone.encoder2_2", 2 * d, 4 * d)
    resnet("backbone.encoder2_3", 4 * d, 4 * d)
    resnet("backbone.encoder3_1", 4 * d, 4 * d)
    resnet("backbone.encoder3_2", 4 * d, 8 * d)
    resnet("backbone.encoder3_3", 8 * d, 8 * d)
    resnet("backbone.encoder4_1", 8 * d, 8 * d)
    resnet("backbone.encoder4_2", 8 * d, 16 * d)
    resnet("backbone.encoder4_3", 16 * d, 16 * d)
    unary("backbone.decoder3", 24 * d, 8 * d)
    shapes["backbone.decoder2.mlp.weight"] = (S["output_dim"], 12 * d)
    shapes["backbone.decoder2.mlp.bias"] = (S["output_dim"],)
    sd = {k: helpers.seeded_tensor(k, s) for k, s in shapes.items()}
    radius = S["init_radius"]
    unit = torch.from_numpy(oe.unit_kernel_points()).float()
    for name in list(sd):
        if name.endswith("interso3.conv.weights"):
            stage = int(name.split("encoder")[1][0])
            block = int(name.split("encoder")[1][2])
            level = stage - 1 if (stage == 1 or block > 1) else stage - 2
            sd[name.replace("weights", "kernel_points")] = unit * (0.7 * radius * 2 ** level)
    return sd


def test_backbone_matches_reference(gold, pyramid):
    S = helpers.SMALL_CFG
    sd = backbone_state_dict()
    feats = torch.ones(gold["in_points"].shape[0], 1)
    out = oe.e2pn_forward(sd, feats, pyramid, S["init_sigma"], S["group_norm"])
    for got, want in zip(out, (gold["feats_f"], gold["feats_mid"], gold["feats_c"])):
        want = torch.from_numpy(want)
        assert got.shape == want.shape
        assert torch.allclose(got, want, rtol=2e-3, atol=2e-4), (got - want).abs().max()
