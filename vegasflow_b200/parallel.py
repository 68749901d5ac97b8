"""
Multi-GPU plumbing: one process per GPU (torchrun), events sharded statically,
one all-reduce of the packed [n_dim*50 + 2] buffer per iteration.

Replaces the reference's joblib device pool and host-side `_accumulate`
(src/vegasflow/monte_carlo.py:143-157, 318-365, 454-480, 72-92): rank r of R
evaluates the global event indices [r*N/R, (r+1)*N/R) of the same Philox
stream, so the union over ranks is exactly the single-GPU event set and every
rank refines an identical grid from the identical reduced histogram.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_events, rank=None, world_size=None):
    """Global event range [begin, end) owned by `rank`."""
    if rank is None or world_size is None:
        rank, world_size = world()
    begin = (int(n_events) * rank) // world_size
    end = (int(n_events) * (rank + 1)) // world_size
    return begin, end


def allreduce_sum_(packed):
    """In-place SUM all-reduce of the packed per-iteration buffer (no-op for one rank)."""
    if world()[1] > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return packed
