"""Metropolis-Hastings sampler driving the native MH kernels.

Same dataclass, state tuple and method signatures as src/nqmc/sampler/metropolis.py:23-231.
``log_prob_fn`` must be the bound closure ``wavefunction.log_prob_fn(params)``; walkers stay
resident on the GPU for the whole of ``sample`` (burn-in + sampling) and only the kept samples
are copied back, instead of materialising every intermediate position (metropolis.py:202-213).

Extensions (keyword-only, default off): ``normals`` / ``uniforms`` to ``step`` supply the proposal
and acceptance streams explicitly -- the parity mode of SURVEY A.11.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, NamedTuple, Optional, Tuple

import numpy as np

from .. import random as nqrandom
from ..wavefunction.base import require_native, ravel


class SamplerState(NamedTuple):
    positions: np.ndarray      # (n_chains, n_electrons, 3) float32
    log_prob: np.ndarray       # (n_chains,) float32
    n_accepted: np.ndarray     # (n_chains,) float32
    n_steps: int


@dataclass
class MetropolisSampler:
    step_size: float = 0.1
    target_acceptance: float = 0.5      # declared, never read -- as in the reference (metropolis.py:53-54)
    adapt_step: bool = True

    # ---------------------------------------------------------------------------------------
    def init_state(self, key, log_prob_fn: Callable, n_chains: int, n_electrons: int, init_width: float = 0.5,
                   init_positions=None) -> SamplerState:
        bound = require_native(log_prob_fn)
        if init_positions is None:
            init_positions = np.zeros((n_electrons, 3), dtype=np.float32)
        init_positions = np.asarray(init_positions, dtype=np.float32)
        positions = init_positions[None] + np.float32(init_width) * nqrandom.normal(key, (n_chains, n_electrons, 3))
        log_prob = np.asarray(bound(positions), dtype=np.float32)            # metropolis.py:90
        return SamplerState(positions.astype(np.float32), log_prob, np.zeros(n_chains, dtype=np.float32), 0)

    # ---------------------------------------------------------------------------------------
    def step(self, key, state: SamplerState, log_prob_fn: Callable, *, normals=None, uniforms=None) -> SamplerState:
        import torch
        bound = require_native(log_prob_fn)
        ctx = bound.wavefunction.context()
        ctx.set_params(ravel(bound.params))
        r = ctx.to_soa(state.positions)
        dev = ctx.dev()
        logp = torch.as_tensor(np.asarray(state.log_prob, dtype=np.float32)).to(dev)
        nacc = torch.as_tensor(np.asarray(state.n_accepted, dtype=np.float32)).to(dev)
        n_dev = ctx.to_soa(normals) if normals is not None else None
        u_dev = torch.as_tensor(np.asarray(uniforms, dtype=np.float32)).to(dev) if uniforms is not None else None
        ctx.mh_step(r, logp, self.step_size, normals=n_dev, uniforms=u_dev, seed=nqrandom.key_to_seed(key), step=0,
                    n_accepted=nacc)
        return SamplerState(ctx.to_aos(r).cpu().numpy(), logp.cpu().numpy(), nacc.cpu().numpy(), state.n_steps + 1)

    # ---------------------------------------------------------------------------------------
    def sample(self, key, log_prob_fn: Callable, state: SamplerState, n_samples: int, n_burn: int = 100,
               n_thin: int = 1) -> Tuple[np.ndarray, SamplerState, float]:
        samples_dev, final, rate, _ = self.sample_device(key, log_prob_fn, state, n_samples, n_burn, n_thin)
        ctx = require_native(log_prob_fn).wavefunction.context()
        n_chains = state.positions.shape[0]
        # samples_dev: SoA [3N, n_samples * n_chains] with flat index s * n_chains + c (vmc.py:105 order)
        aos = ctx.to_aos(samples_dev).cpu().numpy().reshape(n_samples, n_chains, ctx.n_elec, 3)
        return aos, final, rate

    def sample_device(self, key, log_prob_fn: Callable, state: SamplerState, n_samples: int, n_burn: int, n_thin: int,
                      walker_id0: int = 0):
        """Device-resident body of ``sample``: returns (samples SoA tensor, final state, rate, ctx)."""
        import torch
        bound = require_native(log_prob_fn)
        ctx = bound.wavefunction.context()
        ctx.set_params(ravel(bound.params))
        dev = ctx.dev()
        W = state.positions.shape[0]
        r = ctx.to_soa(state.positions)
        logp = torch.as_tensor(np.asarray(state.log_prob, dtype=np.float32)).to(dev)
        nacc = torch.zeros(W, dtype=torch.float32, device=dev)                # reset (metropolis.py:177-182)
        k_burn, k_samp = nqrandom.split(key, 2)
        seed_b, seed_s = nqrandom.key_to_seed(k_burn), nqrandom.key_to_seed(k_samp)
        for t in range(n_burn):
            ctx.mh_step(r, logp, self.step_size, seed=seed_b, step=t, walker_id0=walker_id0, n_accepted=nacc)
        nacc.zero_()                                                          # reset again (metropolis.py:194-199)
        n_total = n_samples * n_thin
        out = torch.empty((3 * ctx.n_elec, n_samples, W), dtype=torch.float32, device=dev)
        for t in range(n_total):
            ctx.mh_step(r, logp, self.step_size, seed=seed_s, step=t, walker_id0=walker_id0, n_accepted=nacc)
            if t % n_thin == 0:                                               # all_positions[::n_thin] (:213)
                out[:, t // n_thin, :].copy_(r)
        rate = float(nacc.mean().item()) / max(n_total, 1)                    # metropolis.py:216
        final = SamplerState(ctx.to_aos(r).cpu().numpy(), logp.cpu().numpy(), nacc.cpu().numpy(), n_total)
        return out.reshape(3 * ctx.n_elec, n_samples * W), final, rate, ctx

    def get_acceptance_rate(self, state: SamplerState) -> float:
        if state.n_steps == 0:
            return 0.0
        return float(np.mean(state.n_accepted) / state.n_steps)
