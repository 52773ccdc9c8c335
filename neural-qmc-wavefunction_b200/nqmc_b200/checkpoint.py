"""Checkpoint / resume of a VMC run.

The reference only sketches ``save_checkpoint(path, params, energy)`` (.github/phase4-issue-body.md:177-183) and
reserves ``checkpoints/ *.ckpt *.pkl *.npz`` in its .gitignore; ``VMCState`` (src/nqmc/optimizer/vmc.py:33-46) is a
plain tuple of arrays, so a checkpoint here is one ``.npz``: theta in ``jax.tree.leaves`` order, the Adam moments and
step count, the walker positions / log-probabilities / acceptance counters, and the VMC step number.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .params import Params, ravel, unravel
from .sampler.metropolis import SamplerState

FORMAT_VERSION = 2


def rank_path(path: str) -> str:
    """Walkers are sharded over ranks, so every rank writes its own file: ``name.npz`` for a single process,
    ``name.rank{r}of{w}.npz`` otherwise."""
    from . import parallel
    rank, ws = parallel.world()
    base = str(path)[:-4] if str(path).endswith(".npz") else str(path)
    return base + (f".rank{rank}of{ws}" if ws > 1 else "") + ".npz"


def save_checkpoint(path: str, state, energy: Optional[float] = None, key=None) -> None:
    """state: nqmc_b200.optimizer.VMCState.  Device-resident walkers / Adam moments are copied back here."""
    from . import parallel
    rank, ws = parallel.world()
    opt = state.opt_state
    mu, nu = np.asarray(opt["mu"]), np.asarray(opt["nu"])
    ss = state.sampler_state
    np.savez(rank_path(path), format_version=FORMAT_VERSION, rank=rank, world_size=ws, theta=ravel(state.params), opt_count=int(opt["count"]),
             opt_mu=mu.astype(np.float32), opt_nu=nu.astype(np.float32),
             positions=np.asarray(ss.positions, np.float32), log_prob=np.asarray(ss.log_prob, np.float32),
             n_accepted=np.asarray(ss.n_accepted, np.float32), n_steps=int(ss.n_steps), step=int(state.step),
             energy=np.float64(np.nan if energy is None else energy),
             key=np.zeros(2, np.uint32) if key is None else np.asarray(key, np.uint32))


def load_checkpoint(path: str, like_params: Params):
    """-> (VMCState, energy, key).  ``like_params`` supplies the tree structure (e.g. ``wf.init(key)``)."""
    from .optimizer.vmc import VMCState
    from . import parallel
    z = np.load(rank_path(path))
    if int(z["format_version"]) != FORMAT_VERSION:
        raise ValueError(f"unsupported checkpoint version {int(z['format_version'])}")
    rank, ws = parallel.world()
    if (int(z["rank"]), int(z["world_size"])) != (rank, ws):
        raise ValueError(f"checkpoint was written by rank {int(z['rank'])} of {int(z['world_size'])}, "
                         f"this process is rank {rank} of {ws}")
    params = unravel(z["theta"], like_params)
    opt = {"count": int(z["opt_count"]), "mu": z["opt_mu"].copy(), "nu": z["opt_nu"].copy()}
    ss = SamplerState(z["positions"].copy(), z["log_prob"].copy(), z["n_accepted"].copy(), int(z["n_steps"]))
    return VMCState(params, opt, ss, int(z["step"])), float(z["energy"]), z["key"].copy()
