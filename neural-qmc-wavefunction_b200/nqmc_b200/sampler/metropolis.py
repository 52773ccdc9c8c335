"""Metropolis-Hastings sampler driving the native MH kernels.

Same dataclass, state tuple and method signatures as src/nqmc/sampler/metropolis.py:23-231.
``log_prob_fn`` must be the bound closure ``wavefunction.log_prob_fn(params)``; walkers stay resident on the GPU:
the arrays inside a ``SamplerState`` returned from here are device buffers (``WalkerArray`` / ``DeviceArray``) that
convert to NumPy on demand (``np.asarray(state.positions)`` gives the reference's ``(n_chains, n_electrons, 3)``), so
threading a state through ``sample`` / ``VMCOptimizer.step`` never round-trips through the host.  Discarded
positions are never materialised (the reference stacks all of them and slices, metropolis.py:202-213).

Extensions (keyword-only, default off):
* ``normals`` / ``uniforms`` to ``step`` / ``sample`` supply the proposal and acceptance streams explicitly --
  the parity mode of SURVEY A.11;
* ``adaptive=True`` makes ``target_acceptance`` / ``adapt_step`` -- which the reference declares and never reads
  (metropolis.py:52-54) -- do what their docstring says: the proposal width is adapted during burn-in only.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, NamedTuple, Optional, Tuple

import numpy as np

from .. import random as nqrandom
from .._native import DeviceArray
from ..wavefunction.base import require_native, ravel


class SamplerState(NamedTuple):
    positions: object          # (n_chains, n_electrons, 3) float32: ndarray or WalkerArray (device-resident)
    log_prob: object           # (n_chains,) float32: ndarray or DeviceArray
    n_accepted: object         # (n_chains,) float32
    n_steps: int


class WalkerArray:
    """Walker positions resident on the GPU in the library's SoA layout [3N][W]; looks like the reference's
    (W, N, 3) array to NumPy (``np.asarray``, ``.shape``), converting at the edge only when asked."""

    def __init__(self, ctx, soa: DeviceArray):
        self.ctx, self.soa = ctx, soa

    @property
    def shape(self):
        return (self.soa.shape[1], self.ctx.n_elec, 3)

    dtype = np.dtype(np.float32)
    ndim = 3

    def __len__(self):
        return self.soa.shape[1]

    def numpy(self) -> np.ndarray:
        return self.ctx.to_aos(self.soa).numpy(self.ctx.stream)

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype)

    def clone(self) -> "WalkerArray":
        return WalkerArray(self.ctx, self.soa.clone(self.ctx.stream))


def _resident(ctx, state: SamplerState, donate: bool):
    """-> (r_soa, logp) device buffers for ``state``; fresh copies unless ``donate`` (then updated in place)."""
    p = state.positions
    if isinstance(p, WalkerArray) and p.ctx is ctx:
        r = p.soa if donate else p.soa.clone(ctx.stream)
    else:
        r = ctx.to_soa(np.asarray(p, dtype=np.float32))
    lp = state.log_prob
    if isinstance(lp, DeviceArray) and lp.device == ctx.device:
        logp = lp if donate else lp.clone(ctx.stream)
    else:
        logp = ctx.upload(np.asarray(lp, dtype=np.float32))
    return r, logp


def _streams_to_device(ctx, normals, uniforms, n_steps: int, W: int):
    """explicit streams (steps, W, N, 3) / (steps, W) -> device SoA [steps][3N][W] / [steps][W]"""
    n_dev = u_dev = None
    if normals is not None:
        normals = np.asarray(normals, dtype=np.float32)
        if normals.shape != (n_steps, W, ctx.n_elec, 3):
            raise ValueError(f"normals must have shape {(n_steps, W, ctx.n_elec, 3)}, got {normals.shape}")
        n_dev = ctx.empty((n_steps, 3 * ctx.n_elec, W))
        for t in range(n_steps):
            a = ctx.upload(normals[t])
            ctx._check(ctx.lib.nqmc_aos_to_soa(ctx.h, a.ptr, n_dev.row(t).ptr, W, ctx.stream), "nqmc_aos_to_soa")
            ctx.synchronize()
    if uniforms is not None:
        uniforms = np.asarray(uniforms, dtype=np.float32)
        if uniforms.shape != (n_steps, W):
            raise ValueError(f"uniforms must have shape {(n_steps, W)}, got {uniforms.shape}")
        u_dev = ctx.upload(uniforms)
    return n_dev, u_dev


@dataclass
class MetropolisSampler:
    step_size: float = 0.1
    target_acceptance: float = 0.5      # read only when adaptive=True (the reference never reads it, metropolis.py:53-54)
    adapt_step: bool = True
    adaptive: bool = False              # extension: adapt step_size toward target_acceptance during burn-in
    adapt_every: int = 10               # burn-in moves per adaptation block
    _sigma_dev: object = field(default=None, repr=False, compare=False)

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
        """One MH move for every chain (metropolis.py:99-147)."""
        bound = require_native(log_prob_fn)
        ctx = bound.wavefunction.context()
        ctx.set_params(ravel(bound.params))
        r, logp = _resident(ctx, state, donate=False)
        nacc = ctx.upload(np.asarray(state.n_accepted, dtype=np.float32))
        W = r.shape[1]
        n_dev, u_dev = _streams_to_device(ctx, None if normals is None else np.asarray(normals)[None],
                                          None if uniforms is None else np.asarray(uniforms)[None], 1, W)
        ctx.mh_step(r, logp, self.step_size, normals=n_dev, uniforms=u_dev, seed=nqrandom.key_to_seed(key), step=0,
                    n_accepted=nacc)
        return SamplerState(WalkerArray(ctx, r), logp, nacc, state.n_steps + 1)

    # ---------------------------------------------------------------------------------------
    def sample(self, key, log_prob_fn: Callable, state: SamplerState, n_samples: int, n_burn: int = 100,
               n_thin: int = 1, *, normals=None, uniforms=None) -> Tuple[np.ndarray, SamplerState, float]:
        """(samples (n_samples, n_chains, N, 3), final state, acceptance rate) -- metropolis.py:149-218."""
        samples_dev, final, rate, ctx = self.sample_device(key, log_prob_fn, state, n_samples, n_burn, n_thin,
                                                           normals=normals, uniforms=uniforms)
        n_chains = final.positions.shape[0]
        # samples_dev: SoA [3N, n_samples * n_chains] with flat index s * n_chains + c (vmc.py:105 order)
        aos = ctx.to_aos(samples_dev).numpy(ctx.stream).reshape(n_samples, n_chains, ctx.n_elec, 3)
        return aos, final, rate

    def sample_device(self, key, log_prob_fn: Callable, state: SamplerState, n_samples: int, n_burn: int, n_thin: int,
                      walker_id0: int = 0, *, donate: bool = False, normals=None, uniforms=None, want_rate: bool = True):
        """Device-resident body of ``sample``: returns (samples SoA [3N, n_samples*W] DeviceArray, final state,
        rate, ctx).  One library call (nqmc_sample); nothing touches the host except the acceptance rate."""
        bound = require_native(log_prob_fn)
        ctx = bound.wavefunction.context()
        ctx.set_params(ravel(bound.params))
        r, logp = _resident(ctx, state, donate)
        W = r.shape[1]
        nacc = ctx.empty(W)
        k_burn, k_samp = nqrandom.split(key, 2)
        n_total = n_samples * n_thin
        n_dev, u_dev = _streams_to_device(ctx, normals, uniforms, n_burn + n_total, W)
        out = ctx.empty((3 * ctx.n_elec, max(n_samples, 1) * W))
        sig = None
        if self.adaptive and self.adapt_step:
            if self._sigma_dev is None or self._sigma_dev.device != ctx.device:
                self._sigma_dev = ctx.upload(np.array([self.step_size], dtype=np.float32))
            sig = self._sigma_dev
        ctx.sample(r, logp, self.step_size, n_burn, n_samples, n_thin, out, n_accepted=nacc,
                   seed_burn=nqrandom.key_to_seed(k_burn), seed_sample=nqrandom.key_to_seed(k_samp),
                   walker_id0=walker_id0, normals=n_dev, uniforms=u_dev, sigma_dev=sig,
                   target_acceptance=self.target_acceptance, adapt_every=self.adapt_every if sig is not None else 0)
        if sig is not None:
            self.step_size = float(sig.numpy(ctx.stream)[0])
        rate = float(nacc.numpy(ctx.stream).mean()) / max(n_total, 1) if want_rate else float("nan")   # metropolis.py:216
        final = SamplerState(WalkerArray(ctx, r), logp, nacc, n_total)
        return out, final, rate, ctx

    def get_acceptance_rate(self, state: SamplerState) -> float:
        if state.n_steps == 0:
            return 0.0
        return float(np.mean(np.asarray(state.n_accepted)) / state.n_steps)
