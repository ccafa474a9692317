"""GPU parity of the E2PN backbone kernels/modules against the torch-fp32 oracle (oracle/e2pn.py) and the
reference fixtures.  Tolerances (bf16 tensor-core operands, fp32 accumulation):
  single op, bf16-rounded inputs on both sides: rtol 2e-2 / atol 2e-3 (scaled to the output magnitude);
  whole backbone: cosine similarity >= 0.999 per point on feats_c / feats_f and rel. Frobenius error <= 2e-2."""
import os

import numpy as np
import pytest
import torch

import helpers
from oracle import e2pn as oe
from oracle import points as op
from se3et_b200.modules import e2pn as M
from se3et_b200.modules import octahedral
from se3et_b200.ops import e2pn_ops as K
from test_oracle_e2pn import backbone_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class EpnCfg:
    num_kernel_points, kanchor, quotient_factor = 15, 6, 4
    KP_influence, aggregation_mode = 'linear', 'sum'
    epn_kernel, equiv_mode_kp, non_sep_conv, rot_by_permute = False, True, True, True
    fixed_kernel_points, ignore_steer_constraint, gather_by_idxing = 'center', False, False
    batch_norm_momentum, att_pooling, att_permute = 0.99, False, False


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "model_small.npz"))


@pytest.fixture(scope="module")
def pyramid(gold):
    S = helpers.SMALL_CFG
    return op.precompute_data_stack_mode(gold["in_points"], gold["in_lengths"], 4, S["init_voxel"], S["init_radius"],
                                         [38, 36, 36, 38], impl="oracle")


def rel_err(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def test_tables_match_kernel_and_reference(gold):
    t = octahedral.tables()
    kidx, ridx = K.builtin_tables()
    assert np.array_equal(kidx, t["kidx"]) and np.array_equal(ridx, t["ridx"])
    assert np.array_equal(t["kidx"], gold["const_kidx_rot"][:, 0, :])
    assert np.array_equal(t["ridx"], gold["const_ridx_rot"][0])
    assert np.allclose(t["anchors"], gold["const_anchors"], atol=1e-6)
    conv = M.KPConvInterSO3(15, 6, 8, 16, 0.05, 0.0625, non_sep_conv=True, rot_by_permute=True, quotient_factor=4)
    assert np.allclose(conv.kernel_points.numpy(), gold["const_kernel_points"], atol=1e-7)
    assert set(conv.state_dict().keys()) == {"kernel_points", "quotient_anchors", "anchors", "weights", "kidx_rot",
                                             "ridx_rot"}
    assert tuple(conv.kidx_rot.shape) == (15, 6, 6) and tuple(conv.ridx_rot.shape) == (15, 6, 6)


@pytest.mark.parametrize("cin,cout", [(8, 16), (1, 16), (32, 32), (64, 128)])
def test_kpconv_matches_oracle(pyramid, cin, cout):
    t = oe.octahedral_tables()
    p1 = torch.from_numpy(pyramid["points"][1])
    p0 = torch.from_numpy(pyramid["points"][0])
    for q, s, nb in ((p1, p1, pyramid["neighbors"][1]), (p1, p0, pyramid["subsampling"][0])):
        nb = torch.from_numpy(nb)
        conv = M.KPConvInterSO3(15, 6, cin, cout, 0.05, 0.0625, non_sep_conv=True, rot_by_permute=True,
                                quotient_factor=4)
        with torch.no_grad():
            conv.weights.copy_(helpers.seeded_tensor("conv.weights", (6, 6, cin, cout)))
        x = helpers.seeded_tensor("conv.input", (s.shape[0], 6, cin)).bfloat16().float()
        w = conv.weights.detach().bfloat16().float()
        want = oe.kpconv_inter_so3(q, s, nb, x, w, conv.kernel_points.detach(), 0.05, t["kidx"], t["ridx"])
        conv = conv.to(DEV)
        got = conv(q.to(DEV), s.to(DEV), nb.to(DEV), x.to(DEV)).cpu()
        assert got.shape == want.shape
        # the gathered operand is rounded to bf16 once more before the GEMM: atol scales with the output magnitude
        assert torch.allclose(got, want, rtol=2e-2, atol=5e-3 * want.abs().max().item() + 1e-6), \
            (got - want).abs().max().item()
        assert rel_err(got, want) < 5e-3


def test_kpconv_matches_reference_fixture(gold, pyramid):
    p1 = torch.from_numpy(pyramid["points"][1]).to(DEV)
    nb1 = torch.from_numpy(pyramid["neighbors"][1]).to(DEV)
    conv = M.KPConvInterSO3(15, 6, 8, 16, 0.05, 0.0625, non_sep_conv=True, rot_by_permute=True, quotient_factor=4)
    with torch.no_grad():
        conv.weights.copy_(helpers.seeded_tensor("conv.weights", (6, 6, 8, 16)))
    x = helpers.seeded_tensor("conv.input", (p1.shape[0], 6, 8))
    got = conv.to(DEV)(p1, p1, nb1, x.to(DEV)).cpu()
    want = torch.from_numpy(gold["conv_out"])
    assert rel_err(got, want) < 1e-2  # inputs AND weights rounded to bf16 here


def test_groupnorm_pairs_do_not_mix():
    g = torch.Generator().manual_seed(3)
    n1, n2, c, G = 700, 513, 32, 4
    y = torch.randn(n1 + n2, 6, c, generator=g) * 2 + 0.5
    gamma, beta = torch.randn(c, generator=g), torch.randn(c, generator=g)
    want = torch.cat([oe.group_norm_epn(y[:n1], G, gamma, beta), oe.group_norm_epn(y[n1:], G, gamma, beta)])
    want = torch.nn.functional.leaky_relu(want, 0.1)
    seg = torch.tensor([0, n1, n1 + n2], dtype=torch.int64, device=DEV)
    yd = y.reshape(-1, c).to(DEV)
    stats = K.groupnorm_stats(yd, G, seg, 6)
    of, ob = K.groupnorm_apply(yd, stats, gamma.to(DEV), beta.to(DEV), G, seg, 6, slope=0.1, out_f32=True)
    assert torch.allclose(of.cpu().view_as(want), want, rtol=1e-4, atol=1e-4)
    assert torch.allclose(ob.float().cpu().view_as(want), want, rtol=1e-2, atol=1e-2)


def test_pooling_ops_match_oracle(pyramid):
    g = torch.Generator().manual_seed(4)
    n0, n1 = pyramid["points"][0].shape[0], pyramid["points"][1].shape[0]
    x = torch.randn(n0, 6, 16, generator=g).bfloat16()
    sub = torch.from_numpy(pyramid["subsampling"][0])
    want = oe.max_pool(x.float(), sub)
    got = K.maxpool_nbr(x.to(DEV), sub.to(DEV)).float().cpu()
    assert torch.equal(got, want)
    assert torch.equal(K.anchor_max(x.to(DEV)).float().cpu(), x.float().amax(1))
    lat = torch.randn(n1, 24, generator=g).bfloat16()
    skip = torch.randn(n0, 8, generator=g).bfloat16()
    up = torch.from_numpy(pyramid["upsampling"][0])
    want = torch.cat([oe.nearest_upsample(lat.float(), up), skip.float()], 1)
    got = K.upsample_concat(lat.to(DEV), up.to(DEV), skip.to(DEV)).float().cpu()
    assert torch.equal(got, want)


def _build_backbone():
    S = helpers.SMALL_CFG
    net = M.E2PN(S["input_dim"], S["output_dim"], S["init_dim"], S["init_radius"], S["init_sigma"], S["group_norm"],
                 EpnCfg)
    sd = backbone_state_dict()
    own = net.state_dict()
    for k, v in sd.items():
        name = k[len("backbone."):]
        assert name in own, name
        assert own[name].shape == v.shape, name
    missing, unexpected = net.load_state_dict({k[len("backbone."):]: v for k, v in sd.items()}, strict=False)
    assert not unexpected
    assert all(helpers.is_constant(m) for m in missing), missing
    return net.to(DEV).eval(), sd


def test_backbone_matches_oracle_and_reference(gold, pyramid):
    S = helpers.SMALL_CFG
    net, sd = _build_backbone()
    dd = {k: [torch.from_numpy(np.ascontiguousarray(a)).to(DEV) for a in pyramid[k]]
          for k in ("points", "neighbors", "subsampling", "upsampling")}
    feats = torch.ones(gold["in_points"].shape[0], 1, device=DEV)
    with torch.no_grad():
        out = net(feats, dd)
    want = oe.e2pn_forward(sd, feats.cpu(), pyramid, S["init_sigma"], S["group_norm"])
    for name, g, w, ref in zip(("feats_f", "feats_mid", "feats_c"), out, want,
                               (gold["feats_f"], gold["feats_mid"], gold["feats_c"])):
        g = g.float().cpu()
        assert g.shape == w.shape, name
        cos = torch.nn.functional.cosine_similarity(g.reshape(g.shape[0], -1), w.reshape(w.shape[0], -1), dim=1)
        assert cos.min().item() > 0.999, (name, cos.min().item())
        assert rel_err(g, w) < 2e-2, (name, rel_err(g, w))
        assert rel_err(g, torch.from_numpy(ref)) < 2e-2, name


def test_backbone_batched_pairs_equal_single_pairs(gold, pyramid):
    """Two pairs stacked in one launch (pair_offsets) give each pair's own result: GroupNorm never mixes pairs."""
    S = helpers.SMALL_CFG
    net, _ = _build_backbone()
    pts_a, lens_a = gold["in_points"], gold["in_lengths"]
    pts_b, lens_b = helpers.small_pair(index=12, crop=1.1)
    from se3et_b200.precompute import precompute_data_stack_mode
    outs = []
    for pts, lens in ((pts_a, lens_a), (pts_b, lens_b)):
        d = precompute_data_stack_mode(torch.from_numpy(pts).to(DEV), torch.from_numpy(lens).to(DEV), 4,
                                       S["init_voxel"], S["init_radius"], [38, 36, 36, 38])
        with torch.no_grad():
            outs.append([o.float() for o in net(torch.ones(len(pts), 1, device=DEV), d)])
    pts = np.concatenate([pts_a, pts_b])
    lens = np.concatenate([lens_a, lens_b])
    d = precompute_data_stack_mode(torch.from_numpy(pts).to(DEV), torch.from_numpy(lens).to(DEV), 4, S["init_voxel"],
                                   S["init_radius"], [38, 36, 36, 38])
    assert [o.tolist() for o in d["pair_offsets"]] == [[0, int(l[:2].sum()), int(l.sum())] for l in d["lengths"]]
    with torch.no_grad():
        both = [o.float() for o in net(torch.ones(len(pts), 1, device=DEV), d)]
    for lvl, (o, a, b) in zip((1, 2, 3), zip(both, outs[0], outs[1])):
        na = a.shape[0]
        assert o.shape[0] == na + b.shape[0]
        ea, eb = rel_err(o[:na], a), rel_err(o[na:], b)
        # GroupNorm statistics are accumulated in fp64, so they do not depend on how rows are split over CTAs;
        # what is left is fp64 summation-order noise (1e-16) that can flip a bf16 rounding once in a blue moon
        assert ea < 2e-3 and eb < 2e-3, (lvl, ea, eb)
