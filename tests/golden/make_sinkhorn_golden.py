"""Generates tests/golden/sinkhorn_ref.npz from the UNMODIFIED reference module
geotransformer.modules.sinkhorn.learnable_sinkhorn.LearnableLogOptimalTransport (CPU; `.cuda()` neutralised by the shim).

    python tests/golden/make_sinkhorn_golden.py        (needs /root/reference)
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
from geotransformer.modules.sinkhorn import LearnableLogOptimalTransport  # noqa: E402


def main():
    out = {}
    g = torch.Generator().manual_seed(7)
    for tag, (b, m, n, iters, masked, alpha) in {"small": (6, 16, 12, 100, True, 1.0), "patch": (3, 64, 64, 100, True, 0.37),
                                                 "nomask": (2, 9, 20, 30, False, -0.5)}.items():
        scores = torch.randn(b, m, n, generator=g) * 2.0
        rm = cm = None
        if masked:
            rm = torch.rand(b, m, generator=g) > 0.25
            cm = torch.rand(b, n, generator=g) > 0.25
            rm[0] = True
            cm[0] = True
        ot = LearnableLogOptimalTransport(iters)
        with torch.no_grad():
            ot.alpha.fill_(alpha)
            res = ot(scores, rm, cm)
        out[tag + "_scores"] = scores.numpy()
        out[tag + "_alpha"] = np.float32(alpha)
        out[tag + "_iters"] = np.int64(iters)
        if masked:
            out[tag + "_row_masks"], out[tag + "_col_masks"] = rm.numpy(), cm.numpy()
        out[tag + "_out"] = res.numpy()
        print(tag, tuple(res.shape), float(res[res > -1e11].min()), float(res.max()))
    np.savez_compressed(os.path.join(HERE, "sinkhorn_ref.npz"), **out)


if __name__ == "__main__":
    main()
