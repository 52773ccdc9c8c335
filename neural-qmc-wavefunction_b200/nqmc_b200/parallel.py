"""Multi-GPU plumbing: walker sharding and the per-iteration all-reduces (one process per GPU).

The reference has no collective.  Walkers are independent Markov chains
(src/nqmc/sampler/metropolis.py:44,124); the only coupling is the mean / variance / gradient
reduction of ``compute_gradient`` (src/nqmc/optimizer/vmc.py:113-133).  Rank g owns a contiguous
block of global walker ids, so device RNG streams -- keyed by global id -- do not depend on the
number of ranks.  Two exchanges per iteration, both additive:
   hdr  (8 doubles: n, sum E, sum E^2, n_accept, n_bad)   before the reverse pass
   gsum (P doubles: sum_w (E_w - mean E) O_w)             after it
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) carries them.
"""
from __future__ import annotations

from typing import Tuple


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def shard_range(n_walkers: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of global walker ids owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(int(n_walkers), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(t):
    """In-place sum over ranks of a tensor (no-op for a single process)."""
    rank, ws = world()
    if ws > 1:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def attach_peer_memory(ctx) -> bool:
    """Give ``ctx`` (a native ``Context``) the peer-memory communicator: afterwards ``ctx.walker_step`` /
    ``walker_step_host`` sum ``hdr`` and ``gsum`` over all ranks in-stream with a one-shot all-reduce over NVLink
    (csrc/nqmc_comm.cu) instead of two NCCL calls.  The handles travel through ``torch.distributed``.
    Returns False (and leaves ``ctx`` rank-local) for a single process or a world larger than one node's 8 GPUs."""
    rank, ws = world()
    if ws <= 1 or ws > 8:
        return False
    import torch.distributed as dist
    mine = ctx.comm_create(rank, ws)
    handles = [None] * ws
    dist.all_gather_object(handles, mine)
    ctx.comm_attach(handles)
    dist.barrier()
    return True


def combine_gradient(hdr, gsum):
    """gradient = 2 * sum_w (E_w - mean) O_w / n   (vmc.py:126-133) from all-reduced sums.
    Returns (grad, mean_energy, energy_std) with energy_std = std(ddof=0)/sqrt(n) (vmc.py:114)."""
    n = float(hdr[0])
    mean = float(hdr[1]) / n
    var = max(float(hdr[2]) / n - mean * mean, 0.0)
    return 2.0 * gsum / n, mean, (var ** 0.5) / (n ** 0.5)
