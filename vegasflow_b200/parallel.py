"""
Multi-GPU plumbing: one process per GPU (torchrun), events sharded statically,
one all-reduce of the packed [n_dim*50 + 2] buffer per iteration.

Replaces the reference's joblib device pool and host-side `_accumulate`
(src/vegasflow/monte_carlo.py:143-157, 318-365, 454-480, 72-92): rank r of R
evaluates the global event indices [r*N/R, (r+1)*N/R) of the same Philox
stream, so the union over ranks is exactly the single-GPU event set and every
rank refines an identical grid from the identical reduced histogram.
"""
import ctypes as C
import logging
import os

import torch
import torch.distributed as dist

logger = logging.getLogger(__name__)


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


def cube_shard(ev_offset, rank=None, world_size=None):
    """VEGAS+ over several GPUs: cube range [c_lo, c_hi) owned by `rank`.

    `ev_offset` is the exclusive prefix sum of n_ev (n_cubes + 1 entries, events are ordered by
    cube, vflowplus.py:67).  Rank r owns the cubes from the first one starting at or after event
    n*r/R up to the first one starting at or after n*(r+1)/R: contiguous, balanced on the event
    count to within one cube, and a pure function of the offsets -- every rank derives the same
    partition without communicating.  Host-side mirror of `first_cube_at_or_after` in
    csrc/vf_common.cuh (the kernels compute it on the device); used by the CPU tests.
    """
    import numpy as np

    if rank is None or world_size is None:
        rank, world_size = world()
    off = np.asarray(ev_offset, dtype=np.int64)
    n = int(off[-1])
    lo = int(np.searchsorted(off, n * rank // world_size, side="left"))
    hi = int(np.searchsorted(off, n * (rank + 1) // world_size, side="left"))
    return lo, hi


def allreduce_sum_(packed):
    """In-place SUM all-reduce of the packed per-iteration buffer (no-op for one rank)."""
    if world()[1] > 1:
        dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    return packed


class PeerExchange:
    """Symmetric (peer-mapped) exchange buffer for the fused reduce + NVLink all-reduce +
    refine kernel (`vf_run_iteration_sharded`).  One per integrator instance.

    The buffer is allocated with torch symmetric memory so every rank holds device pointers to
    all peers' copies; the kernel pushes its [n_dim*50+2] sums (VEGAS+: also the per-cube
    variances of its cube range) into every peer with 16-byte P2P stores that carry the value AND
    the exchange sequence number, so arrival is detected on the data itself -- no fence, no
    separate flag, no NCCL call on the iteration path.
    """

    def __init__(self, n_dim, device, n_cubes=0):
        import torch.distributed._symmetric_memory as symm_mem

        from vegasflow_b200 import _lib

        rank, world_size = world()
        if world_size > 8:
            raise RuntimeError("the peer exchange covers the GPUs of one NVLink box (<= 8)")
        nbytes = _lib.load().vf_exchange_bytes(int(n_dim), world_size, int(n_cubes))
        self.buf = symm_mem.empty(nbytes // 8, dtype=torch.int64, device=device)
        self.hdl = symm_mem.rendezvous(self.buf, dist.group.WORLD)
        self.buf.zero_()
        torch.cuda.synchronize(device)
        dist.barrier()
        ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        if len(ptrs) != world_size:
            raise RuntimeError("symmetric memory rendezvous returned a wrong number of peers")
        self.ptrs = (C.c_uint64 * world_size)(*ptrs)
        self.rank, self.world_size = rank, world_size
        self.seq = 0

    def next_seq(self):
        self.seq += 1
        return self.seq

    def check(self):
        """Raise when an exchange on this buffer timed out on any rank (the kernels then poison
        every rank's buffer and return NaN instead of hanging the GPU)."""
        if int(self.buf[-1].item()) != 0:
            raise RuntimeError(
                "vegasflow_b200: a peer-memory exchange timed out (a rank is missing, crashed, or "
                "ran a different number of iterations); results of this instance are invalid. "
                "VEGASFLOW_B200_EXCHANGE_TIMEOUT_S sets the bound.")


def make_peer_exchange(n_dim, device, n_cubes=0):
    """PeerExchange, or None when it is disabled (VEGASFLOW_B200_EXCHANGE=nccl) or the platform
    cannot provide peer-mapped memory (then the NCCL all-reduce path is used)."""
    if world()[1] <= 1 or os.environ.get("VEGASFLOW_B200_EXCHANGE", "p2p").lower() == "nccl":
        return None
    try:
        return PeerExchange(n_dim, device, n_cubes)
    except Exception as exc:  # pylint: disable=broad-except
        logger.warning("peer-memory exchange unavailable (%s); using the NCCL all-reduce", exc)
        return None
