"""Generates tests/golden/model_small.npz by running the UNMODIFIED reference python modules (imported from
/root/reference through tests/golden/ref_import_shim.py) on a small seeded problem:

  * group tables / kernel points of KPConvInterSO3        (blocks_epn.py:111-332)
  * one KPConvInterSO3.forward                            (blocks_epn.py:454-546)
  * E2PN backbone forward, reduced width (init_dim 16)    (experiments/se3eti.3dmatch/backbone.py)
  * GeometricTransformer forward (SE3ET-I block list)     (geotransformer.py:213-317)
  * SuperPointMatching forward                            (superpoint_matching.py:13-55)

Weights come from tests/helpers.seeded_state_dict (deterministic from parameter names), so the fixture holds only
inputs and outputs.   python tests/golden/make_model_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import ref_import_shim as shim  # noqa: E402

cfg = shim.make_cfg("se3eti.3dmatch")
import helpers  # noqa: E402
from oracle import points as op  # noqa: E402
from backbone import E2PN  # noqa: E402  (reference experiments/se3eti.3dmatch/backbone.py)
from geotransformer.modules.e2pn.blocks_epn import KPConvInterSO3  # noqa: E402
from geotransformer.modules.geotransformer import GeometricTransformer, SuperPointMatching  # noqa: E402


def load(module, sd, prefix):
    """Loads the seeded values; group-table buffers are expanded views in the reference and stay as built."""
    own = {k[len(prefix):]: v for k, v in sd.items() if not helpers.is_constant(k)}
    missing, unexpected = module.load_state_dict(own, strict=False)
    assert not unexpected and all(helpers.is_constant(k) for k in missing), (missing, unexpected)


def main():
    S = helpers.SMALL_CFG
    out = {}
    torch.manual_seed(0)
    pts, lens = helpers.small_pair()
    d = op.precompute_data_stack_mode(pts, lens, 4, S["init_voxel"], S["init_radius"], [38, 36, 36, 38], impl="oracle")
    out["in_points"], out["in_lengths"] = pts, lens

    # ---- constants + single conv
    conv = KPConvInterSO3(cfg.epn.num_kernel_points, cfg.epn.kanchor, 8, 16, S["init_sigma"], S["init_radius"],
                          cfg.epn.KP_influence, cfg.epn.aggregation_mode, epn_kernel=False,
                          equiv_mode_kp=cfg.epn.equiv_mode_kp, non_sep_conv=True, rot_by_permute=True,
                          fixed_kernel_points='center', quotient_factor=4, ignore_steer_constraint=False,
                          gather_by_idxing=False)
    out["const_kernel_points"] = conv.kernel_points.detach().numpy()
    out["const_anchors"] = conv.anchors.detach().numpy()
    out["const_quotient_anchors"] = conv.quotient_anchors.detach().numpy()
    out["const_kidx_rot"] = conv.kidx_rot.numpy()
    out["const_ridx_rot"] = conv.ridx_rot.numpy()
    sd = helpers.seeded_state_dict({"conv." + k: v for k, v in conv.state_dict().items()})
    load(conv, sd, "conv.")
    p1 = torch.from_numpy(d["points"][1])
    nb1 = torch.from_numpy(d["neighbors"][1])
    x = helpers.seeded_tensor("conv.input", (p1.shape[0], 6, 8))
    with torch.no_grad():
        out["conv_out"] = conv(p1, p1, nb1, x).numpy()

    # ---- backbone
    backbone = E2PN(S["input_dim"], S["output_dim"], S["init_dim"], S["init_radius"], S["init_sigma"], S["group_norm"],
                    cfg.epn)
    bsd = helpers.seeded_state_dict({"backbone." + k: v for k, v in backbone.state_dict().items()})
    load(backbone, bsd, "backbone.")
    backbone.eval()
    data_dict = {k: [torch.from_numpy(np.ascontiguousarray(a)) for a in d[k]]
                 for k in ("points", "neighbors", "subsampling", "upsampling")}
    feats = torch.ones(pts.shape[0], 1)
    with torch.no_grad():
        feats_list = backbone(feats, data_dict)
    out["feats_f"] = feats_list[0].numpy()
    out["feats_mid"] = feats_list[1].numpy()
    out["feats_c"] = feats_list[2].numpy()

    # ---- transformer (SE3ET-I block list) on the coarse level
    nc = d["lengths"][3]
    ref_pc = torch.from_numpy(d["points"][3][:nc[0]])
    src_pc = torch.from_numpy(d["points"][3][nc[0]:])
    tr = GeometricTransformer(16 * S["init_dim"], S["tr_output_dim"], S["hidden_dim"], S["num_heads"], S["blocks"],
                              S["sigma_d"], S["sigma_a"], S["angle_k"], supervise_rotation=False, reduction_a='max',
                              na=6, attn_r_positive='softplus', attn_r_positive_rot_supervise='minus', align_mode='0',
                              alternative_impl=False, n_level_equiv=0)
    tsd = helpers.seeded_state_dict({"transformer." + k: v for k, v in tr.state_dict().items()})
    load(tr, tsd, "transformer.")
    tr.eval()
    fc = feats_list[2]
    with torch.no_grad():
        ref_e = tr.embedding(ref_pc[None])
        rf, sf, _, _, _, _ = tr(ref_pc[None], src_pc[None], fc[:nc[0]][None], fc[nc[0]:][None])
    out["ref_embedding"] = ref_e[0].numpy().astype(np.float16)
    out["ref_feats_c"], out["src_feats_c"] = rf[0].numpy(), sf[0].numpy()

    # ---- superpoint matching
    spm = SuperPointMatching(64, True)
    rn = torch.nn.functional.normalize(rf[0], p=2, dim=1)
    sn = torch.nn.functional.normalize(sf[0], p=2, dim=1)
    rmask = torch.ones(rn.shape[0], dtype=torch.bool)
    smask = torch.ones(sn.shape[0], dtype=torch.bool)
    rmask[3] = False
    smask[[0, 7]] = False
    with torch.no_grad():
        ri, si, sc = spm(rn, sn, rmask, smask)
    out["spm_ref_feats"], out["spm_src_feats"] = rn.numpy(), sn.numpy()
    out["spm_ref_masks"], out["spm_src_masks"] = rmask.numpy(), smask.numpy()
    out["spm_ref_idx"], out["spm_src_idx"], out["spm_scores"] = ri.numpy(), si.numpy(), sc.numpy()

    path = os.path.join(HERE, "model_small.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()})
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
