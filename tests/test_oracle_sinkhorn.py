"""CPU: the numpy restatement of LearnableLogOptimalTransport (oracle/sinkhorn.py) against the outputs of the reference
module itself (tests/golden/sinkhorn_ref.npz, made by tests/golden/make_sinkhorn_golden.py)."""
import os

import numpy as np
import pytest

from oracle import sinkhorn as osk

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sinkhorn_ref.npz")


def load(tag):
    g = np.load(GOLD)
    rm = g[tag + "_row_masks"] if tag + "_row_masks" in g.files else None
    cm = g[tag + "_col_masks"] if tag + "_col_masks" in g.files else None
    return g[tag + "_scores"], float(g[tag + "_alpha"]), int(g[tag + "_iters"]), rm, cm, g[tag + "_out"]


@pytest.mark.parametrize("tag", ["small", "patch", "nomask"])
def test_oracle_matches_reference_sinkhorn(tag):
    scores, alpha, iters, rm, cm, want = load(tag)
    got = osk.log_optimal_transport(scores, alpha, iters, rm, cm)
    assert got.shape == want.shape
    live = want > -1e11  # masked entries sit at ~ -1e12, where fp32 spacing is 65536
    assert np.array_equal(live, got > -1e11)
    assert np.abs(got[live] - want[live]).max() < 2e-4
    assert np.allclose(got[~live], want[~live], rtol=1e-6)
    # the transport plan's marginals: rows / columns of exp(out + norm) sum to mu / nu
    if rm is None:
        b, m, n = scores.shape
        p = np.exp(got.astype(np.float64) - np.log(m + n))
        assert np.allclose(p[:, :m].sum(2), 1.0 / (m + n), rtol=1e-3)
