"""Host-side pieces of the training step (se3et_b200/training.py) against the UNMODIFIED reference
(tests/golden/loss_ref.npz from tests/golden/make_loss_golden.py): ground-truth superpoint correspondences, weighted
circle loss, fine matching loss, the ATen log-domain optimal transport used for the backward pass (values and
gradients), and the one-bucket gradient exchange over a 2-rank gloo group."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from se3et_b200 import training as TR


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "loss_ref.npz"))


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def test_node_correspondences_match_reference(gold):
    idx, ov = TR.get_node_correspondences(*[t(gold["nc_" + k]) for k in ("ref_nodes", "src_nodes", "ref_knn", "src_knn", "T")],
                                          0.05, *[t(gold["nc_" + k]) for k in ("ref_masks", "src_masks", "ref_km", "src_km")])
    assert np.array_equal(idx.numpy(), gold["nc_gt_idx"])
    assert np.allclose(ov.numpy(), gold["nc_gt_ov"], rtol=1e-6)


def test_coarse_loss_and_gradient_match_reference(gold):
    r, s = t(gold["cl_ref"]).requires_grad_(True), t(gold["cl_src"]).requires_grad_(True)
    loss = TR.coarse_matching_loss(r, s, t(gold["nc_gt_idx"]), t(gold["nc_gt_ov"]))
    loss.backward()
    assert abs(float(loss) - float(gold["cl_loss"])) < 1e-5
    assert np.allclose(r.grad.numpy(), gold["cl_gref"], rtol=1e-4, atol=1e-6)
    assert np.allclose(s.grad.numpy(), gold["cl_gsrc"], rtol=1e-4, atol=1e-6)


def test_optimal_transport_and_fine_loss_match_reference(gold):
    scores = t(gold["fl_scores"]).requires_grad_(True)
    alpha = torch.tensor(0.4, requires_grad=True)
    b = scores.shape[0]
    rm, cm = t(gold["nc_ref_km"])[:b], t(gold["nc_src_km"])[:b]
    ms = TR.aten_log_optimal_transport(scores, alpha, 20, rm, cm)
    live = gold["fl_ms"] > -1e11
    assert np.array_equal(live, ms.detach().numpy() > -1e11)
    assert np.abs(ms.detach().numpy()[live] - gold["fl_ms"][live]).max() < 1e-4
    loss = TR.fine_matching_loss(ms, t(gold["nc_ref_knn"])[:b], t(gold["nc_src_knn"])[:b], rm, cm, t(gold["nc_T"]))
    loss.backward()
    assert abs(float(loss) - float(gold["fl_loss"])) < 1e-4
    assert np.allclose(scores.grad.numpy(), gold["fl_gscores"], rtol=1e-3, atol=1e-6)
    assert abs(float(alpha.grad) - float(gold["fl_galpha"])) < 1e-4


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2, 2))]
    params[0].grad = torch.full((5, 3), float(rank + 1))
    params[1].grad = None                                   # a parameter that received no gradient on this rank
    params[2].grad = torch.arange(4.0).view(2, 2) * (rank + 1)
    nbytes = TR.allreduce_gradients(params, world)
    out[rank] = (nbytes, params[0].grad.clone(), params[1].grad.clone(), params[2].grad.clone())
    dist.destroy_process_group()


def test_gradient_bucket_allreduce_two_ranks_gloo():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    for rank in (0, 1):
        nbytes, g0, g1, g2 = out[rank]
        assert nbytes == (15 + 7 + 4) * 4
        assert torch.allclose(g0, torch.full((5, 3), 1.5)) and torch.allclose(g1, torch.zeros(7))
        assert torch.allclose(g2, torch.arange(4.0).view(2, 2) * 1.5)
