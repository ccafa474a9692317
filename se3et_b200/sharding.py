"""Multi-GPU plumbing of the inference path: pairs are independent, so the path shards by pair with no data-path
collective (the reference runs batch size 1 per process, experiments/se3eti.3dmatch/config.py:39,48).
One process per GPU; pair i goes to rank i mod world.  torch.distributed is used only for the barrier and for
reducing the per-rank device time to its maximum (bench.py) -- NCCL on GPUs, gloo in the CPU tests."""
import torch
import torch.distributed as dist


def pairs_for_rank(num_pairs, rank, world):
    """Indices of the pairs rank `rank` of `world` processes owns (round robin)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, num_pairs, world))


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_results(local, device=None):
    """All ranks' python result lists, flattened back into global pair order: `local` = [(pair_index, payload), ...]
    for the pairs this rank owns.  Used after the timed region only (results are host objects)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [p for _, p in sorted(local, key=lambda kv: kv[0])]
    world = dist.get_world_size()
    bucket = [None] * world
    dist.all_gather_object(bucket, local)
    merged = [kv for part in bucket for kv in part]
    idx = [k for k, _ in merged]
    if sorted(idx) != list(range(len(idx))):
        raise RuntimeError("pair shards do not partition the batch: %r" % sorted(idx))
    return [p for _, p in sorted(merged, key=lambda kv: kv[0])]
