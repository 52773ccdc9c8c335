"""Stochastic Reconfiguration on the device.

Mirror of the reference's [SPEC] ``StochasticReconfiguration`` (``.github/phase6-issue-body.md:130-182``; the
reference ships no ``optimizer/sr.py``): same constructor arguments, same ``compute_update(log_psi_grads,
local_energies)`` signature and return value (``-lr * solve(S, f)``).  The work happens in ``libnqmc_b200.so``:
``S = O^T O / n`` is a TMA-fed tcgen05 3xTF32 SYRK, the solve is preconditioned CG on the device (``nqmc_sr_update``).
"""
from __future__ import annotations

import numpy as np

from .. import _native


class StochasticReconfiguration:
    """delta_theta = -lr * S^-1 * dE/dtheta with S the covariance of the log-derivatives (phase6-issue-body.md:133-147)."""

    def __init__(self, learning_rate: float = 1e-2, regularization: float = 1e-4, max_iter: int = 500, tol: float = 1e-7,
                 context: "_native.Context | None" = None):
        self.lr = learning_rate
        self.reg = regularization
        self.max_iter = max_iter
        self.tol = tol
        self._ctx = context
        self.last_info = None

    def _context(self):
        if self._ctx is None:
            # a bare context (hydrogen atom, SimpleWavefunction shape) only to own the stream / device for the SR kernels
            self._ctx = _native.Context(np.zeros((1, 3)), np.ones(1), 1, 0, 0.0, kind=_native.KIND_SIMPLE, hidden=(8, 8))
        return self._ctx

    def compute_update(self, log_psi_grads, local_energies):
        """log_psi_grads: (n_samples, n_params) host array, or a DeviceArray already PARAMETER-major (n_params, n_samples);
        local_energies: (n_samples,).  Returns the update direction as float64 NumPy (n_params,)."""
        ctx = self._context()
        if isinstance(log_psi_grads, _native.DeviceArray):
            Ot = log_psi_grads
            n = Ot.shape[1]
            if n % 4:
                raise ValueError("device log_psi_grads need a sample count that is a multiple of 4")
        else:
            g = np.asarray(log_psi_grads, dtype=np.float32)
            if g.ndim != 2:
                raise ValueError("log_psi_grads must be (n_samples, n_params)")
            n = g.shape[0]
            ld = (n + 3) // 4 * 4
            gt = np.zeros((g.shape[1], ld), dtype=np.float32)
            gt[:, :n] = g.T
            Ot = ctx.upload(gt)
        e = np.asarray(local_energies, dtype=np.float32).reshape(-1) if not isinstance(local_energies, _native.DeviceArray) else None
        if e is not None:
            if e.shape[0] != n:
                raise ValueError("local_energies and log_psi_grads disagree on n_samples")
            ee = np.full(Ot.shape[1], np.nan, dtype=np.float32)     # padding samples: non-finite energy -> dropped
            ee[:n] = e
            e_dev = ctx.upload(ee)
        else:
            e_dev = local_energies
        delta, info, _ = ctx.sr_update(Ot, e_dev, self.reg, self.lr, self.max_iter, self.tol)
        self.last_info = dict(zip(("iterations", "relative_residual", "converged", "f_norm"), info.numpy(ctx.stream).tolist()))
        return delta.numpy(ctx.stream)
