"""TEST INFRASTRUCTURE (called by __graft_entry__.smoke() only): one tiny end-to-end invocation of the hot path,
SE3ET-I2 on a small synthetic pair -- pyramid + backbone + transformer + SuperPointMatching on the GPU -- checked against the
torch-CPU oracle.  Lives under oracle/ because it imports the oracle; nothing in se3et_b200/ does."""
import numpy as np
import torch


def run(dev):
    from oracle import e2pn as oe
    from oracle import points as op
    from oracle import transformer as ot
    from se3et_b200 import synthetic
    from se3et_b200.model import create_model, make_cfg

    cfg = make_cfg("se3eti2.3dmatch")
    torch.manual_seed(0)
    model = create_model(cfg)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(dev).eval()
    p = synthetic.make_3dmatch_pair(13, crop=0.9)
    ref, src = p["ref_points"], p["src_points"]
    got = model.forward_pairs([(ref, src)])[0]
    b, g = cfg.backbone, cfg.geotransformer
    pts, lens = np.concatenate([ref, src]), np.array([len(ref), len(src)])
    d = op.precompute_data_stack_mode(pts, lens, b.num_stages, b.init_voxel_size, b.init_radius, cfg.neighbor_limits,
                                      impl="oracle")
    with torch.no_grad():
        fl = oe.e2pn_forward(sd, torch.ones(len(pts), 1), d, b.init_sigma, b.group_norm)
        n = int(d["lengths"][-1][0])
        pc = torch.from_numpy(d["points"][-1])
        r, s, _, _ = ot.geometric_transformer(sd, pc[:n], pc[n:], fl[-1][:n], fl[-1][n:], g.blocks, g.hidden_dim,
                                              g.num_heads, g.sigma_d, g.sigma_a, g.angle_k)
        r = torch.nn.functional.normalize(r, p=2, dim=1)
        s = torch.nn.functional.normalize(s, p=2, dim=1)
        ri, si, _ = ot.superpoint_matching(r, s, torch.ones(len(r), dtype=torch.bool), torch.ones(len(s), dtype=torch.bool),
                                           cfg.coarse_matching.num_correspondences)
    want = set(zip(ri.tolist(), si.tolist()))
    have = set(zip(got[0].tolist(), got[1].tolist()))
    overlap = len(want & have) / max(1, len(want))
    assert overlap >= 0.85, "coarse correspondences differ from the oracle: overlap %.2f" % overlap
    print("smoke: SE3ET-I2 forward on %d + %d points, %d correspondences, overlap with the oracle %.2f"
          % (len(ref), len(src), len(have), overlap))
