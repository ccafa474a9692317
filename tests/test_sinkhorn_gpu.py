"""GPU: se3et_log_optimal_transport (through the C ABI and the LearnableLogOptimalTransport mirror) against the numpy
oracle and the reference fixtures; fp32, tolerance 2e-4 absolute on the log-domain scores (reduction orders differ)."""
import os

import numpy as np
import pytest
import torch

from oracle import sinkhorn as osk

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sinkhorn_ref.npz")


def run(scores, alpha, iters, rm, cm):
    from se3et_b200.modules.sinkhorn import LearnableLogOptimalTransport
    ot = LearnableLogOptimalTransport(iters).to(DEV)
    assert list(ot.state_dict().keys()) == ["alpha"]
    with torch.no_grad():
        ot.alpha.fill_(alpha)
    t = lambda x: None if x is None else torch.from_numpy(x).to(DEV)
    return ot(t(scores), t(rm), t(cm)).cpu().numpy()


@pytest.mark.parametrize("tag", ["small", "patch", "nomask"])
def test_sinkhorn_matches_reference_fixture(tag):
    g = np.load(GOLD)
    rm = g[tag + "_row_masks"] if tag + "_row_masks" in g.files else None
    cm = g[tag + "_col_masks"] if tag + "_col_masks" in g.files else None
    got = run(g[tag + "_scores"], float(g[tag + "_alpha"]), int(g[tag + "_iters"]), rm, cm)
    want = g[tag + "_out"]
    live = want > -1e11
    assert np.array_equal(live, got > -1e11)
    assert np.abs(got[live] - want[live]).max() < 2e-4
    assert np.allclose(got[~live], want[~live], rtol=1e-6)


@pytest.mark.parametrize("b,m,n,iters", [(5, 64, 64, 100), (3, 128, 128, 100), (4, 1, 7, 10), (2, 33, 200, 50), (0, 8, 8, 5)])
def test_sinkhorn_matches_oracle(b, m, n, iters):
    rng = np.random.default_rng(m * 1000 + n)
    scores = (rng.standard_normal((b, m, n)) * 3).astype(np.float32)
    rm = rng.random((b, m)) > 0.3
    cm = rng.random((b, n)) > 0.3
    if b:
        rm[:, 0] = True
        cm[:, 0] = True
    got = run(scores, 0.8, iters, rm, cm)
    assert got.shape == (b, m + 1, n + 1)
    if b == 0:
        return
    want = osk.log_optimal_transport(scores, 0.8, iters, rm, cm)
    live = want > -1e11
    assert np.array_equal(live, got > -1e11)
    assert np.abs(got[live] - want[live]).max() < 3e-4


def test_sinkhorn_full_size_marginals():
    """256 patch pairs x 32 pairs of 64 x 64 scores (the fine stage of one launch sequence): the marginals of the plan."""
    from se3et_b200.modules.sinkhorn import log_optimal_transport
    b, k = 8192, 64
    g = torch.Generator(device="cpu").manual_seed(0)
    scores = (torch.randn(b, k, k, generator=g) * 2).to(DEV)
    out = log_optimal_transport(scores, torch.tensor(1.0, device=DEV), 100)
    p = torch.exp(out.double() - np.log(2 * k))
    assert torch.allclose(p[:, :k].sum(2), torch.full((b, k), 1.0 / (2 * k), dtype=torch.float64, device=DEV), rtol=2e-3)
    assert torch.allclose(p[:, :, :k].sum(1), torch.full((b, k), 1.0 / (2 * k), dtype=torch.float64, device=DEV), rtol=2e-2)
