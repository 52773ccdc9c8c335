"""Wavefunction interface (host side) bound to the native walker kernels.

Keeps the reference's contract -- ``wf(params, r) -> log|psi|``, ``wf.init(key, r_sample)``,
``wf.log_prob(params, r) = 2 log|psi|`` (src/nqmc/wavefunction/base.py:38-76) -- but a call is a
batched kernel launch through libnqmc_b200, not a traced Python function.  ``r`` may be one
configuration ``(N,3)`` (the reference's signature) or a batch ``(W,N,3)``.

Because the reference differentiates a per-walker closure with jax.grad / hessian / vmap, a kernel
hidden behind ``__call__`` alone would still be driven by autodiff.  The sampler, Hamiltonian and
optimizer of this package therefore recognise these classes and dispatch whole batches
(SURVEY section 8(b)); arbitrary Python callables are rejected rather than run on the CPU.
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Dict, Optional, Tuple

import numpy as np

from .. import _native
from ..params import Params, ravel


class BoundLogProb:
    """``log_prob_fn`` closure of the reference (optimizer/vmc.py:238-239) made inspectable."""

    __nqmc_native__ = True

    def __init__(self, wavefunction: "Wavefunction", params: Params):
        self.wavefunction, self.params = wavefunction, params

    def __call__(self, r):
        return self.wavefunction.log_prob(self.params, r)


class Wavefunction(ABC):
    kind: int = -1

    # ---- geometry / context ---------------------------------------------------------------
    def _geometry(self, molecule=None) -> Tuple[np.ndarray, np.ndarray, int, int, float]:
        """(R, Z, n_up, n_dn, v_nn) from an explicit molecule or from the constructor arguments."""
        raise NotImplementedError

    def _net_args(self) -> Dict[str, Any]:
        raise NotImplementedError

    def context(self, molecule=None) -> "_native.Context":
        R, Z, n_up, n_dn, v_nn = self._geometry(molecule)
        key = (R.tobytes(), Z.tobytes(), n_up, n_dn, float(v_nn))
        cache = self.__dict__.setdefault("_ctx_cache", {})
        if key not in cache:
            cache[key] = _native.Context(R, Z, n_up, n_dn, v_nn, kind=self.kind, **self._net_args())
        return cache[key]

    # ---- reference API ----------------------------------------------------------------------
    @abstractmethod
    def init(self, key, r_sample) -> Params:
        ...

    def sign_and_log(self, params: Params, r, molecule=None):
        """(sign, log|psi|); the [SPEC] flax modules return this pair (phase3-issue-body.md:219)."""
        ctx = self.context(molecule)
        ctx.set_params(ravel(params))
        r = np.asarray(r, dtype=np.float32)
        single = r.ndim == 2
        if not single and r.shape[0] == 0:                  # empty batch
            return np.zeros((0,), np.float32), np.zeros((0,), np.float32)
        soa = ctx.to_soa(r[None] if single else r)
        sg, la = ctx.log_psi(soa)
        sg, la = sg.cpu().numpy(), la.cpu().numpy()
        return (sg[0], la[0]) if single else (sg, la)

    def __call__(self, params: Params, r):
        return self.sign_and_log(params, r)[1]

    def log_prob(self, params: Params, r):
        return 2.0 * self(params, r)                      # base.py:76

    def log_prob_fn(self, params: Params) -> BoundLogProb:
        return BoundLogProb(self, params)


def require_native(log_prob_fn) -> BoundLogProb:
    if not getattr(log_prob_fn, "__nqmc_native__", False):
        raise TypeError(
            "log_prob_fn must come from Wavefunction.log_prob_fn(params): the B200 walker loop runs whole batches "
            "through CUDA kernels and has no CPU path for arbitrary Python closures")
    return log_prob_fn


def lecun_normal(rng: np.random.Generator, shape) -> np.ndarray:
    """flax ``nn.Dense`` default kernel init: truncated normal (+-2 sigma), variance 1/fan_in."""
    n = int(np.prod(shape))
    x = rng.standard_normal(2 * n + 16)
    x = x[np.abs(x) <= 2.0][:n]
    return (x / 0.87962566103423978 / np.sqrt(shape[0])).reshape(shape).astype(np.float32)


def dense_params(rng, n_in: int, n_out: int) -> Dict[str, np.ndarray]:
    return {"kernel": lecun_normal(rng, (n_in, n_out)), "bias": np.zeros(n_out, dtype=np.float32)}
