"""N > 1 host logic on CPU: two gloo ranks shard a batch of pairs round-robin, each "processes" its shard, results are
gathered back in pair order and the per-rank time is reduced to its maximum (what bench.py does around the GPU path)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from se3et_b200 import sharding


def test_round_robin_partition():
    for n in (0, 1, 7, 64):
        for world in (1, 2, 4, 8):
            parts = [sharding.pairs_for_rank(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        sharding.pairs_for_rank(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, num_pairs, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.pairs_for_rank(num_pairs, rank, world)
        local = [(i, {"pair": i, "rank": rank, "checksum": i * i + 1}) for i in mine]
        t = sharding.max_over_ranks(10.0 + rank)
        dist.barrier()
        merged = sharding.gather_results(local)
        if rank == 0:
            torch.save({"merged": merged, "tmax": t}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("num_pairs", [5, 8])
def test_two_ranks_gloo(tmp_path, num_pairs):
    world, port = 2, _free_port()
    out = str(tmp_path / "merged.pt")
    mp.spawn(_worker, args=(world, port, num_pairs, out), nprocs=world, join=True)
    res = torch.load(out)
    assert res["tmax"] == 11.0
    assert [m["pair"] for m in res["merged"]] == list(range(num_pairs))
    assert [m["rank"] for m in res["merged"]] == [i % world for i in range(num_pairs)]
    assert all(m["checksum"] == m["pair"] ** 2 + 1 for m in res["merged"])
