"""Generates tests/golden/model_e_small.npz: the UNMODIFIED reference GeometricTransformer with the SE3ET-E block list
(experiments/se3ete.3dmatch/config.py:194, shortened to one invariant self/cross pair) on the coarse level of the
small seeded problem of make_model_golden.py.

e3nn is not installed here; the reference calls o3.spherical_harmonics / o3.Irrep.D_from_matrix for the equivariant
embedding (geotransformer.py:52-66).  This script installs a stand-in with the convention stated in
oracle/transformer.py:sh_equiv_embedding, so the fixture pins everything in the SE3ET-E path EXCEPT that convention.
    python tests/golden/make_model_e_golden.py
"""
import math
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
o3 = sys.modules["e3nn.o3"]


def _spherical_harmonics(ls, x, normalize=True, normalization="integral"):
    assert list(ls) == [0, 1] and normalize and normalization == "integral"
    unit = torch.nn.functional.normalize(x, dim=-1)
    y0 = torch.full(x.shape[:-1] + (1,), 0.5 / math.sqrt(math.pi), dtype=x.dtype)
    return torch.cat([y0, math.sqrt(3.0 / (4.0 * math.pi)) * unit], dim=-1)


class _Irrep:
    def __init__(self, l, p):
        self.l = l

    def D_from_matrix(self, R):
        return torch.ones(R.shape[0], 1, 1) if self.l == 0 else R.clone()


o3.spherical_harmonics = _spherical_harmonics
o3.Irrep = _Irrep

import helpers  # noqa: E402
from oracle import points as op  # noqa: E402
import geotransformer.modules.geotransformer.geotransformer as gmod  # noqa: E402
gmod.o3 = o3
from geotransformer.modules.geotransformer import GeometricTransformer  # noqa: E402

BLOCKS_E = ['self_eq', 'cross_a_soft', 'self_eq', 'cross_r_soft', 'self', 'cross']


def main():
    S = helpers.SMALL_CFG
    base = np.load(os.path.join(HERE, "model_small.npz"))
    d = op.precompute_data_stack_mode(base["in_points"], base["in_lengths"], 4, S["init_voxel"], S["init_radius"],
                                      [38, 36, 36, 38], impl="oracle")
    nc = d["lengths"][3]
    ref_pc = torch.from_numpy(d["points"][3][:nc[0]])
    src_pc = torch.from_numpy(d["points"][3][nc[0]:])
    fc = torch.from_numpy(base["feats_c"])
    out = {}
    for tag, nlev in (("sh", 2), ("nosh", 0)):
        tr = GeometricTransformer(16 * S["init_dim"], S["tr_output_dim"], S["hidden_dim"], S["num_heads"], BLOCKS_E,
                                  S["sigma_d"], S["sigma_a"], S["angle_k"], supervise_rotation=False, reduction_a='max',
                                  na=6, align_mode='0', alternative_impl=False, n_level_equiv=nlev)
        tsd = helpers.seeded_state_dict({"transformer_e." + k: v for k, v in tr.state_dict().items()})
        own = {k[len("transformer_e."):]: v for k, v in tsd.items()
               if not helpers.is_constant(k) and "anchors_wignerD" not in k}  # group constants stay as built
        missing, unexpected = tr.load_state_dict(own, strict=False)
        assert not unexpected, unexpected
        tr.eval()
        with torch.no_grad():
            rf, sf, _, _, _, _ = tr(ref_pc[None], src_pc[None], fc[:nc[0]][None], fc[nc[0]:][None])
        out["ref_feats_%s" % tag], out["src_feats_%s" % tag] = rf[0].numpy(), sf[0].numpy()
        if nlev:
            out["param_shapes"] = np.array(["%s:%s" % (k, ",".join(map(str, v.shape))) for k, v in tr.state_dict().items()])
            att = tr.transformer.layers[3].attention.attention
            out["trace_idx_ori"] = att.trace_idx_ori.numpy()
            out["anchors_embedding"] = tr.embedding.anchors_wignerD[1].numpy()
    path = os.path.join(HERE, "model_e_small.npz")
    np.savez_compressed(path, **out)
    print({k: v.shape for k, v in out.items()})
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
