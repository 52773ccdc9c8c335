"""Multi-GPU plumbing: walker sharding, process groups and the per-iteration sums (one process per GPU).

The reference has no collective.  Walkers are independent Markov chains
(src/nqmc/sampler/metropolis.py:44,124); the only coupling is the mean / variance / gradient
reduction of ``compute_gradient`` (src/nqmc/optimizer/vmc.py:113-133).  Rank g owns a contiguous
block of global walker ids, so device RNG streams -- keyed by global id -- do not depend on the
number of ranks.  Two exchanges per iteration, both additive:
   hdr  (8 doubles: n, sum E, sum E^2, n_accept, n_bad)   before the reverse pass
   gsum (P doubles: sum_w (E_w - mean E) O_w)             after it
On GPUs they run in-stream through the library's peer-memory all-reduce (csrc/nqmc_comm.cu) once
``attach_peer_memory`` has exchanged the IPC handles.  The handle exchange, and the sums in CPU tests,
go through a *process group*:

* ``FileGroup``  -- rendezvous through a directory (one node, no dependency); what ``init_from_env`` builds from
  RANK / WORLD_SIZE / MASTER_PORT as set by ``torchrun`` or ``mpirun``;
* ``torch.distributed`` -- used only if the CALLER has already imported and initialised it (gloo in the CPU tests,
  NCCL in bench.py); this module never imports torch.
"""
from __future__ import annotations

import os
import pickle
import sys
import time
from typing import List, Optional, Tuple

import numpy as np

_GROUP = None


class FileGroup:
    """Process group over a shared directory: ``exchange(obj)`` returns every rank's object in rank order."""

    def __init__(self, rank: int, world: int, root: str, timeout: float = 300.0):
        self.rank, self.world, self.root, self.timeout = int(rank), int(world), root, float(timeout)
        self.round = 0
        os.makedirs(root, exist_ok=True)

    def exchange(self, obj) -> List:
        k = self.round
        self.round += 1
        tmp = os.path.join(self.root, f".x{k}_{self.rank}.tmp")
        with open(tmp, "wb") as f:
            pickle.dump(obj, f)
        os.replace(tmp, os.path.join(self.root, f"x{k}_{self.rank}"))       # atomic publish
        out, t0 = [], time.time()
        for r in range(self.world):
            path = os.path.join(self.root, f"x{k}_{r}")
            while not os.path.exists(path):
                if time.time() - t0 > self.timeout:
                    raise TimeoutError(f"rank {self.rank}: no message from rank {r} in round {k} ({self.root})")
                time.sleep(0.0005)
            with open(path, "rb") as f:
                out.append(pickle.load(f))
        return out

    def barrier(self):
        self.exchange(None)

    def all_reduce_sum(self, a: np.ndarray) -> np.ndarray:
        parts = self.exchange(np.asarray(a))
        tot = parts[0].copy()
        for p in parts[1:]:                                   # rank order: identical result on every rank
            tot = tot + p
        return tot


class _TorchGroup:
    """Adapter over an already-initialised torch.distributed default group (never imports torch itself)."""

    def __init__(self, dist):
        self.dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()

    def exchange(self, obj) -> List:
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out

    def barrier(self):
        self.dist.barrier()

    def all_reduce_sum(self, a: np.ndarray) -> np.ndarray:
        torch = sys.modules["torch"]
        t = torch.as_tensor(np.ascontiguousarray(a))
        if self.dist.get_backend() == "nccl":
            t = t.cuda()
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return t.cpu().numpy()


def set_group(group) -> None:
    """Install a process group explicitly (anything with rank, world, exchange, barrier, all_reduce_sum)."""
    global _GROUP
    _GROUP = group


def init_from_env(root: Optional[str] = None) -> Tuple[int, int]:
    """Build a FileGroup from RANK / WORLD_SIZE (torchrun, mpirun ...).  Ranks of one job share the launcher as parent
    process, so the directory is keyed by it and by MASTER_PORT."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if world > 1:
        root = root or os.path.join(os.environ.get("TMPDIR", "/tmp"),
                                    f"nqmc_b200_rdv_{os.getppid()}_{os.environ.get('MASTER_PORT', '0')}")
        set_group(FileGroup(rank, world, root))
    return rank, world


def group():
    if _GROUP is not None:
        return _GROUP
    dist = sys.modules.get("torch.distributed")
    try:
        if dist is not None and dist.is_available() and dist.is_initialized():
            return _TorchGroup(dist)
    except Exception:
        pass
    return None


def world() -> Tuple[int, int]:
    """(rank, world_size); (0, 1) for a single process."""
    g = group()
    return (g.rank, g.world) if g is not None else (0, 1)


def shard_range(n_walkers: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of global walker ids owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(int(n_walkers), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_sum_(t):
    """In-place sum over ranks (no-op for a single process).  ``t``: numpy array, DeviceArray, or -- when the
    caller runs torch.distributed -- a torch tensor."""
    g = group()
    if g is None or g.world <= 1:
        return t
    if isinstance(t, np.ndarray):
        t[...] = g.all_reduce_sum(t)
        return t
    torch = sys.modules.get("torch")
    if torch is not None and isinstance(t, torch.Tensor):
        if isinstance(g, _TorchGroup):
            g.dist.all_reduce(t, op=g.dist.ReduceOp.SUM)
        else:
            t.copy_(torch.as_tensor(g.all_reduce_sum(t.cpu().numpy())))
        return t
    t.copy_from_host(g.all_reduce_sum(t.numpy()))             # DeviceArray: host round trip (fallback path only)
    return t


def attach_peer_memory(ctx) -> bool:
    """Give ``ctx`` (a native ``Context``) the peer-memory communicator: afterwards ``ctx.walker_step`` /
    ``walker_step_host`` / ``ctx.comm_allreduce`` sum over all ranks in-stream with a one-shot all-reduce over NVLink
    (csrc/nqmc_comm.cu).  The 64-byte IPC handles travel through the process group.
    Returns False (and leaves ``ctx`` rank-local) for a single process or a world larger than one node's 8 GPUs.
    Every rank must make the same library calls in the same order afterwards (include/nqmc_b200.h)."""
    g = group()
    if g is None or g.world <= 1 or g.world > 8:
        return False
    handles = g.exchange(ctx.comm_create(g.rank, g.world))
    ctx.comm_attach(handles)
    g.barrier()
    return True


def reduce_over_ranks_(ctx, vec):
    """Sum a float64 device vector over ranks: in-stream peer-memory all-reduce when attached, process group otherwise."""
    if ctx.comm_active():
        ctx.comm_allreduce(vec)
        return vec
    return all_reduce_sum_(vec)


def combine_gradient(hdr, gsum):
    """gradient = 2 * sum_w (E_w - mean) O_w / n   (vmc.py:126-133) from all-reduced sums.
    Returns (grad, mean_energy, energy_std) with energy_std = std(ddof=0)/sqrt(n) (vmc.py:114).
    n == 0 (every walker hit a node / non-finite E_L) raises instead of dividing by zero."""
    n = float(hdr[0])
    if n <= 0:
        raise FloatingPointError("no walker with a finite local energy in this batch (hdr[N] == 0)")
    mean = float(hdr[1]) / n
    var = max(float(hdr[2]) / n - mean * mean, 0.0)
    return 2.0 * gsum / n, mean, (var ** 0.5) / (n ** 0.5)
