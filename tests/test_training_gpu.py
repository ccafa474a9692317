"""Training step on the GPU (se3et_b200/training.py): CUDA forward + ATen-recompute backward against the torch-CPU
oracle differentiated by autograd on the same small pair; the optimal-transport Function against the ATen iterations;
a few optimizer steps on one pair must reduce the reference's loss."""
import numpy as np
import pytest
import torch

from oracle import e2pn as oe
from oracle import points as op
from oracle import transformer as ot
from se3et_b200 import synthetic
from se3et_b200 import training as TR
from se3et_b200.model import create_model, make_cfg
from se3et_b200.precompute import precompute_data_stack_mode

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def cos(a, b):
    a, b = a.flatten().double().cpu(), b.flatten().double().cpu()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-30))


def test_coarse_path_gradients_match_oracle_autograd():
    cfg = make_cfg("se3eti2.3dmatch")
    torch.manual_seed(0)
    model = create_model(cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).train()
    p = synthetic.make_3dmatch_pair(13, crop=0.7)
    ref, src = p["ref_points"], p["src_points"]
    b, g = cfg.backbone, cfg.geotransformer
    pts, lens = np.concatenate([ref, src]), np.array([len(ref), len(src)])
    # ---- oracle: fp32 CPU forward differentiated by autograd
    watch = ["backbone.encoder1_2.interso3.conv.weights", "backbone.encoder3_2.unary2.mlp.weight",
             "backbone.encoder4_3.interso3.conv.weights", "transformer.transformer.layers.0.attention.attention.proj_p.weight",
             "transformer.transformer.layers.3.attention.attention.proj_v.weight", "transformer.in_proj.weight",
             "backbone.decoder2.mlp.weight", "transformer.embedding.proj_a.weight"]
    for k in watch:
        sd[k].requires_grad_(True)
    d = op.precompute_data_stack_mode(pts, lens, b.num_stages, b.init_voxel_size, b.init_radius, cfg.neighbor_limits,
                                      impl="oracle")
    fl = oe.e2pn_forward(sd, torch.ones(len(pts), 1), d, b.init_sigma, b.group_norm)
    n = int(d["lengths"][-1][0])
    pc = torch.from_numpy(d["points"][-1])
    r, s, _, _ = ot.geometric_transformer(sd, pc[:n], pc[n:], fl[-1][:n], fl[-1][n:], g.blocks, g.hidden_dim, g.num_heads,
                                          g.sigma_d, g.sigma_a, g.angle_k)
    r = torch.nn.functional.normalize(r, p=2, dim=1)
    s = torch.nn.functional.normalize(s, p=2, dim=1)
    # a loss that touches every output: coarse circle loss on synthetic ground truth + energy of the fine features
    gen = torch.Generator().manual_seed(1)
    gt_idx = torch.stack([torch.randint(0, len(r), (40,), generator=gen), torch.randint(0, len(s), (40,), generator=gen)], 1)
    gt_idx = torch.unique(gt_idx, dim=0)
    gt_ov = torch.rand(len(gt_idx), generator=gen) * 0.8 + 0.15

    def loss_of(rc, sc, ff):
        return TR.coarse_matching_loss(rc, sc, gt_idx.to(rc.device), gt_ov.to(rc.device)) + (ff ** 2).mean()
    want_loss = loss_of(r, s, fl[0])
    want = torch.autograd.grad(want_loss, [sd[k] for k in watch])
    # ---- CUDA forward, ATen backward
    dd = precompute_data_stack_mode(torch.from_numpy(pts).to(DEV), torch.from_numpy(lens).to(DEV), b.num_stages,
                                    b.init_voxel_size, b.init_radius, cfg.neighbor_limits)
    params = TR.trainable_parameters(model)
    rc, sc, ff = TR._CoarsePath.apply(model, dd, *params)
    assert cos(rc, r) > 0.999 and cos(ff, fl[0]) > 0.999
    got_loss = loss_of(rc, sc, ff)
    got_loss.backward()
    assert abs(float(got_loss) - float(want_loss)) < 2e-2 * abs(float(want_loss))
    named = dict(model.named_parameters())
    for k, w in zip(watch, want):
        gk = named[k].grad
        assert gk is not None and torch.isfinite(gk).all(), k
        # the upstream gradient is evaluated at the bf16 CUDA outputs, the oracle's at fp32 outputs
        assert cos(gk, w) > 0.98, (k, cos(gk, w))
        assert abs(float(gk.norm()) / float(w.norm()) - 1) < 0.1, k


def test_optimal_transport_function_gradients():
    g = torch.Generator().manual_seed(5)
    scores = (torch.randn(6, 24, 20, generator=g) * 1.5).to(DEV).requires_grad_(True)
    alpha = torch.tensor(0.3, device=DEV, requires_grad=True)
    rm = (torch.rand(6, 24, generator=g) > 0.2).to(DEV)
    cm = (torch.rand(6, 20, generator=g) > 0.2).to(DEV)
    rm[:, 0] = True
    cm[:, 0] = True
    weight = torch.randn(6, 25, 21, generator=g).to(DEV)
    out = TR._OptimalTransport.apply(scores, alpha, 30, rm, cm)
    live = out > -1e11
    (out[live] * weight[live]).sum().backward()
    g_s, g_a = scores.grad.clone(), alpha.grad.clone()
    s2, a2 = scores.detach().clone().requires_grad_(True), alpha.detach().clone().requires_grad_(True)
    ref = TR.aten_log_optimal_transport(s2, a2, 30, rm, cm)
    assert torch.equal(ref > -1e11, live) and (ref[live] - out[live]).abs().max() < 2e-4
    (ref[live] * weight[live]).sum().backward()
    assert torch.allclose(g_s, s2.grad, rtol=1e-4, atol=1e-6) and abs(float(g_a) - float(a2.grad)) < 1e-3 * abs(float(a2.grad)) + 1e-6


def test_training_steps_reduce_the_loss():
    cfg = make_cfg("se3eti2.3dmatch")
    torch.manual_seed(0)
    model = create_model(cfg).to(DEV).train()
    p = synthetic.make_3dmatch_pair(5, crop=0.9)
    opt = torch.optim.Adam(TR.trainable_parameters(model), lr=1e-4)
    rng = np.random.default_rng(0)
    losses = []
    for _ in range(6):
        out = TR.training_step(model, p["ref_points"], p["src_points"], p["transform"], optimizer=opt, rng=rng)
        assert np.isfinite(out["loss"]) and out["grad_bytes"] == 0
        losses.append(out["loss"])
    print("PARITY training losses", [round(l, 4) for l in losses])
    assert min(losses[3:]) < losses[0]
