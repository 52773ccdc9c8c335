"""Local energy E_L = H psi / psi through the native forward-Laplacian kernels.

Signatures follow src/nqmc/hamiltonian/molecular.py:36-197.  The kinetic term is NOT computed by
nested autodiff (dense Hessian + trace, molecular.py:60-64) but by propagating value, gradient and
Laplacian forward through the network analytically (csrc/nqmc_orbital.cu, csrc/nqmc_walker.cu);
results agree to fp32 rounding, not bitwise.
"""
from __future__ import annotations

import numpy as np

from ..params import ravel
from ..wavefunction.base import BoundLogProb, Wavefunction

EPSILON = 1e-10


def _parts(wavefunction: Wavefunction, params, r_batch, molecule):
    ctx = wavefunction.context(molecule)
    ctx.set_params(ravel(params))
    soa = ctx.to_soa(np.asarray(r_batch, dtype=np.float32))
    e, parts = ctx.local_energy(soa, want_parts=True)
    return e.cpu().numpy(), parts.cpu().numpy()


def kinetic_energy(log_psi_fn, r):
    """T = -1/2 (lap log|psi| + |grad log|psi||^2) (molecular.py:36-69) for a *bound* native closure."""
    if not isinstance(log_psi_fn, BoundLogProb):
        raise TypeError("kinetic_energy needs wavefunction.log_prob_fn(params); arbitrary closures have no CPU path here")
    r = np.asarray(r, dtype=np.float32)
    _, parts = _parts(log_psi_fn.wavefunction, log_psi_fn.params, r[None] if r.ndim == 2 else r, None)
    return parts[0][0] if r.ndim == 2 else parts[0]


def electron_nuclear_potential(r, molecule):
    """V_en = -sum Z_A / max(|r_i - R_A|, 1e-10) (molecular.py:72-102).  Host utility; the device path
    computes the same term inside the local-energy kernel."""
    r = np.asarray(r)
    d = np.sqrt(((r[:, None, :] - np.asarray(molecule.positions)[None]) ** 2).sum(-1))
    return -np.sum(np.asarray(molecule.charges)[None, :] / np.maximum(d, EPSILON))


def electron_electron_potential(r):
    """V_ee = sum_{i<j} 1/|r_i - r_j|, 0 for fewer than two electrons (molecular.py:105-135).  Host utility."""
    r = np.asarray(r)
    n = r.shape[0]
    if n < 2:
        return np.float32(0.0)
    iu = np.triu_indices(n, k=1)
    d = np.sqrt(((r[iu[0]] - r[iu[1]]) ** 2).sum(-1))
    return np.sum(1.0 / d)


def local_energy(wavefunction, params, r, molecule):
    """E_L for one configuration r (N,3) (molecular.py:138-175)."""
    return local_energy_batched(wavefunction, params, np.asarray(r)[None], molecule)[0]


def local_energy_batched(wavefunction, params, r_batch, molecule):
    """E_L for a batch (B,N,3) -> (B,) (molecular.py:178-197)."""
    if not isinstance(wavefunction, Wavefunction):
        raise TypeError("local_energy needs a native nqmc_b200 Wavefunction")
    r_batch = np.asarray(r_batch, dtype=np.float32)
    if r_batch.shape[0] == 0:                               # empty batch: vmap over nothing
        return np.zeros((0,), dtype=np.float32)
    ctx = wavefunction.context(molecule)
    ctx.set_params(ravel(params))
    soa = ctx.to_soa(r_batch)
    return ctx.local_energy(soa).cpu().numpy()
