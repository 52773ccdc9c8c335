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

FORMAT_VERSION = 1


def save_checkpoint(path: str, state, energy: Optional[float] = None, key=None) -> None:
    """state: nqmc_b200.optimizer.VMCState."""
    opt = state.opt_state
    mu = opt["mu"].cpu().numpy() if hasattr(opt["mu"], "cpu") else np.asarray(opt["mu"])
    nu = opt["nu"].cpu().numpy() if hasattr(opt["nu"], "cpu") else np.asarray(opt["nu"])
    ss = state.sampler_state
    np.savez(path, format_version=FORMAT_VERSION, theta=ravel(state.params), opt_count=int(opt["count"]),
             opt_mu=mu.astype(np.float32), opt_nu=nu.astype(np.float32),
             positions=np.asarray(ss.positions, np.float32), log_prob=np.asarray(ss.log_prob, np.float32),
             n_accepted=np.asarray(ss.n_accepted, np.float32), n_steps=int(ss.n_steps), step=int(state.step),
             energy=np.float64(np.nan if energy is None else energy),
             key=np.zeros(2, np.uint32) if key is None else np.asarray(key, np.uint32))


def load_checkpoint(path: str, like_params: Params):
    """-> (VMCState, energy, key).  ``like_params`` supplies the tree structure (e.g. ``wf.init(key)``)."""
    from .optimizer.vmc import VMCState
    z = np.load(path if str(path).endswith(".npz") else str(path) + ".npz")
    if int(z["format_version"]) != FORMAT_VERSION:
        raise ValueError(f"unsupported checkpoint version {int(z['format_version'])}")
    params = unravel(z["theta"], like_params)
    opt = {"count": int(z["opt_count"]), "mu": z["opt_mu"].copy(), "nu": z["opt_nu"].copy()}
    ss = SamplerState(z["positions"].copy(), z["log_prob"].copy(), z["n_accepted"].copy(), int(z["n_steps"]))
    return VMCState(params, opt, ss, int(z["step"])), float(z["energy"]), z["key"].copy()
