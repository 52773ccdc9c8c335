"""Variational Monte Carlo optimizer on top of the native walker loop.

Mirrors src/nqmc/optimizer/vmc.py: ``VMCState`` (:33-46), ``estimate_energy`` (:49-79),
``compute_gradient`` (:82-135), ``VMCOptimizer.{init,step,train}`` (:138-327) with the same
dataclass fields, metric keys (energy, energy_std, acceptance_rate, grad_norm) and callback
signature.  Differences, all on purpose:
  * sampling, E_L and the gradient estimator run as CUDA kernels over the whole batch; the
    per-sample gradient matrix (n_total x P, vmc.py:123) is never materialised;
  * the gradient is centred with the batch mean exactly as vmc.py:113-133, via two device phases
    (energies -> mean -> weighted reverse pass) so that multi-GPU runs need only two tiny all-reduces;
  * the Adam + global-norm-clip update (optax.chain(clip_by_global_norm, adam), vmc.py:180-184,
    256-261) is restated in NumPy on the host -- the reference runs it on the host too.
Quirk kept: ``sampler_state.log_prob`` is carried across the parameter update.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Callable, Dict, List, NamedTuple, Optional, Tuple

import numpy as np

from .. import parallel
from .. import random as nqrandom
from .._native import HDR_N, HDR_N_BAD, HDR_SUM_E, HDR_SUM_E2
from ..params import Params, ravel, tree_leaves, unravel
from ..sampler.metropolis import MetropolisSampler, SamplerState


class VMCState(NamedTuple):
    params: Params
    opt_state: Any
    sampler_state: SamplerState
    step: int


def _energies_device(wavefunction, params, samples, molecule):
    """-> (ctx, r_soa, e_l, hdr) for samples (n_samples,n_chains,N,3) host array or a device SoA tensor."""
    ctx = wavefunction.context(molecule)
    ctx.set_params(ravel(params))
    if isinstance(samples, np.ndarray):
        flat = np.asarray(samples, dtype=np.float32).reshape(-1, *samples.shape[-2:])     # vmc.py:105
        r = ctx.to_soa(flat)
    else:
        r = samples
    e = ctx.local_energy(r)
    hdr = ctx.energy_stats(e)
    parallel.all_reduce_sum_(hdr)
    return ctx, r, e, hdr


def estimate_energy(wavefunction, params, samples, molecule) -> Tuple[float, float]:
    """(mean_energy, energy_std) (vmc.py:49-79); energy_std = std(ddof=0)/sqrt(n)."""
    _, _, _, hdr = _energies_device(wavefunction, params, samples, molecule)
    h = hdr.cpu().numpy()
    n = h[HDR_N]
    mean = h[HDR_SUM_E] / n
    var = max(h[HDR_SUM_E2] / n - mean * mean, 0.0)
    return np.float32(mean), np.float32(np.sqrt(var) / np.sqrt(n))


def compute_gradient(wavefunction, params, samples, molecule) -> Tuple[Params, float, float]:
    """(gradient pytree, mean_energy, energy_std) (vmc.py:82-135): 2 <(E_L - <E_L>) d log|psi| / d theta>."""
    import torch
    ctx, r, e, hdr = _energies_device(wavefunction, params, samples, molecule)
    gsum = torch.empty(ctx.P, dtype=torch.float64, device=ctx.dev())
    ctx.walker_step_b(r, e, hdr, gsum)
    parallel.all_reduce_sum_(gsum)
    grad, mean, std = parallel.combine_gradient(hdr.cpu().numpy(), gsum.cpu().numpy())
    return unravel(grad.astype(np.float32), params), np.float32(mean), np.float32(std)


# ------------------------------------------------------------------------------------------------
# optax.chain(clip_by_global_norm(c), adam(lr)) restated (b1=0.9, b2=0.999, eps=1e-8)
def _opt_init(params: Params) -> Dict[str, Any]:
    p = ravel(params)
    return {"count": 0, "mu": np.zeros_like(p), "nu": np.zeros_like(p)}


def _opt_update(grad_flat: np.ndarray, opt_state, lr: float, clip: float):
    g = grad_flat.astype(np.float32)
    norm = np.sqrt(np.sum(g.astype(np.float64) ** 2))
    if norm >= clip:                                    # optax: keep g when ||g|| < max_norm
        g = (g / np.float32(norm)) * np.float32(clip)
    b1, b2, eps = 0.9, 0.999, 1e-8
    count = opt_state["count"] + 1
    mu = b1 * opt_state["mu"] + (1 - b1) * g
    nu = b2 * opt_state["nu"] + (1 - b2) * g * g
    mhat = mu / (1 - b1 ** count)
    vhat = nu / (1 - b2 ** count)
    upd = (-lr * mhat / (np.sqrt(vhat) + eps)).astype(np.float32)
    return upd, {"count": count, "mu": mu.astype(np.float32), "nu": nu.astype(np.float32)}


@dataclass
class VMCOptimizer:
    learning_rate: float = 1e-2
    n_samples: int = 200
    n_chains: int = 32
    n_burn: int = 200
    step_size: float = 0.5
    clip_grad: float = 1.0
    device_update: bool = False     # extension: fused clip + Adam on the GPU, theta stays resident

    def _local_chains(self) -> Tuple[int, int]:
        rank, ws = parallel.world()
        return parallel.shard_range(self.n_chains, rank, ws)

    def init(self, key, wavefunction, params: Params, molecule) -> VMCState:
        opt_state = _opt_init(params)                                         # vmc.py:180-184
        sampler = MetropolisSampler(step_size=self.step_size)
        R = np.asarray(molecule.positions, dtype=np.float32)
        n_el = molecule.n_electrons
        init_positions = np.stack([R[i % molecule.n_atoms] for i in range(n_el)], axis=0)    # vmc.py:199-203
        lo, hi = self._local_chains()
        # draw for all global chains, keep this rank's block: positions do not depend on the rank count
        noise = nqrandom.normal(key, (self.n_chains, n_el, 3))[lo:hi]
        positions = init_positions[None] + np.float32(0.5) * noise            # init_width=0.5 (vmc.py:206)
        wavefunction.context(molecule)                                        # bind geometry before first call
        log_prob = np.asarray(wavefunction.log_prob(params, positions), dtype=np.float32)
        st = SamplerState(positions.astype(np.float32), log_prob, np.zeros(hi - lo, dtype=np.float32), 0)
        return VMCState(params=params, opt_state=opt_state, sampler_state=st, step=0)

    def step(self, key, state: VMCState, wavefunction, molecule) -> Tuple[VMCState, Dict[str, float]]:
        import torch
        params = state.params
        log_prob_fn = wavefunction.log_prob_fn(params)                        # vmc.py:238-239
        sampler = MetropolisSampler(step_size=self.step_size)
        lo, hi = self._local_chains()
        samples, sampler_state, acc = sampler.sample_device(
            key, log_prob_fn, state.sampler_state, n_samples=self.n_samples,
            n_burn=self.n_burn if state.step == 0 else 10, n_thin=1, walker_id0=lo)[:3]       # vmc.py:242-248
        rank, ws = parallel.world()
        if self.device_update:
            new_state, metrics = self._device_update(state, wavefunction, molecule, samples, sampler_state, acc, lo, hi)
            return new_state, metrics
        gradient, mean_energy, energy_std = compute_gradient(wavefunction, params, samples, molecule)
        if ws > 1:
            a = torch.tensor([acc * (hi - lo), float(hi - lo)], dtype=torch.float64,
                             device=samples.device if torch.distributed.get_backend() == "nccl" else "cpu")
            parallel.all_reduce_sum_(a)
            acc = float(a[0] / a[1])
        g = ravel(gradient)
        upd, opt_state = _opt_update(g, state.opt_state, self.learning_rate, self.clip_grad)  # vmc.py:256-259
        new_params = unravel(ravel(params) + upd, params)                     # optax.apply_updates
        grad_norm = float(np.sqrt(sum(float(np.sum(np.asarray(l, dtype=np.float64) ** 2)) for l in tree_leaves(gradient))))
        new_state = VMCState(new_params, opt_state, sampler_state, state.step + 1)
        metrics = {"energy": float(mean_energy), "energy_std": float(energy_std), "acceptance_rate": float(acc),
                   "grad_norm": grad_norm}
        return new_state, metrics

    def _device_update(self, state, wavefunction, molecule, samples, sampler_state, acc, lo, hi):
        """E_L, centred gradient sums, global-norm clip and Adam all on the device (nqmc_adam_step)."""
        import torch
        params = state.params
        ctx, r, e, hdr = _energies_device(wavefunction, params, samples, molecule)
        dev = ctx.dev()
        gsum = torch.empty(ctx.P, dtype=torch.float64, device=dev)
        ctx.walker_step_b(r, e, hdr, gsum)
        parallel.all_reduce_sum_(gsum)
        opt = state.opt_state
        mu = torch.as_tensor(opt["mu"]).to(dev, torch.float32).contiguous()
        nu = torch.as_tensor(opt["nu"]).to(dev, torch.float32).contiguous()
        gn = torch.zeros(1, dtype=torch.float64, device=dev)
        count = int(opt["count"]) + 1
        ctx.adam_step(gsum, hdr, mu, nu, count, self.learning_rate, self.clip_grad, grad_norm=gn)
        new_params = unravel(ctx.get_params(), params)
        h = hdr.cpu().numpy()
        n = h[HDR_N]
        mean = h[HDR_SUM_E] / n
        var = max(h[HDR_SUM_E2] / n - mean * mean, 0.0)
        rank, ws = parallel.world()
        if ws > 1:
            a = torch.tensor([acc * (hi - lo), float(hi - lo)], dtype=torch.float64,
                             device=dev if torch.distributed.get_backend() == "nccl" else "cpu")
            parallel.all_reduce_sum_(a)
            acc = float(a[0] / a[1])
        new_state = VMCState(new_params, {"count": count, "mu": mu, "nu": nu}, sampler_state, state.step + 1)
        metrics = {"energy": float(mean), "energy_std": float(np.sqrt(var) / np.sqrt(n)), "acceptance_rate": float(acc),
                   "grad_norm": float(gn.item())}
        return new_state, metrics

    def train(self, key, wavefunction, params: Params, molecule, n_steps: int, log_every: int = 10,
              callback: Optional[Callable] = None) -> Tuple[Params, List[Dict[str, float]]]:
        key, init_key = nqrandom.split(key)
        state = self.init(init_key, wavefunction, params, molecule)
        history = []
        for i in range(n_steps):
            key, step_key = nqrandom.split(key)
            state, metrics = self.step(step_key, state, wavefunction, molecule)
            history.append(metrics)
            if (i + 1) % log_every == 0 and parallel.world()[0] == 0:
                print(f"Step {i+1:4d}: E = {metrics['energy']:.6f} ± {metrics['energy_std']:.6f} Ha, "
                      f"acc = {metrics['acceptance_rate']:.2%}, |∇| = {metrics['grad_norm']:.4f}")
            if callback is not None:
                callback(i, state, metrics)
        return state.params, history
