"""SimpleWavefunction: MLP(flatten r) + exponential envelope (ansatz kind 0).

Host-side mirror of src/nqmc/wavefunction/simple.py:73-176 -- same constructor, same parameter tree
(``params/Dense_{0,1,2}/{kernel,bias}``) -- evaluated by csrc/nqmc_simple.cu: any tuple of 1..6 hidden widths (each <= 128), 'tanh' or 'silu'
(simple.py:41,56-65).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

from .. import _native
from ..random import _generator
from .base import Params, Wavefunction, dense_params


class SimpleWavefunction(Wavefunction):
    kind = _native.KIND_SIMPLE

    def __init__(self, hidden_dims: Tuple[int, ...] = (64, 64), activation: str = "tanh", envelope_decay: float = 1.0,
                 nuclear_positions=None):
        if activation not in ("tanh", "silu"):
            raise ValueError(f"Unknown activation: {activation}")
        if not 1 <= len(hidden_dims) <= 6 or any(not 1 <= int(h) <= 128 for h in hidden_dims):
            raise NotImplementedError("the native SimpleWavefunction kernel takes 1..6 hidden layers of width 1..128")
        self.hidden_dims = tuple(int(h) for h in hidden_dims)
        self.activation = activation
        self.envelope_decay = float(envelope_decay)
        if nuclear_positions is None:
            nuclear_positions = [[0.0, 0.0, 0.0], [1.4, 0.0, 0.0]]      # simple.py:118 default H2
        self.nuclear_positions = np.asarray(nuclear_positions, dtype=np.float64).reshape(-1, 3)
        self._n_elec = None

    def _geometry(self, molecule=None):
        if molecule is not None:
            R = np.asarray(molecule.positions, dtype=np.float64)
            if R.shape != self.nuclear_positions.shape or not np.allclose(R, self.nuclear_positions):
                raise ValueError("molecule geometry differs from the wavefunction's nuclear_positions")
            n_up, n_dn = molecule.spin
            # bind: the sampler's context() (no molecule) and the energy pass resolve to the same native context
            self._bound = (R, np.asarray(molecule.charges, dtype=np.float64), int(n_up), int(n_dn),
                           float(molecule.nuclear_repulsion_energy()))
            self._n_elec = int(n_up) + int(n_dn)
        if getattr(self, "_bound", None) is not None:
            return self._bound
        if self._n_elec is None:
            raise RuntimeError("call init(key, r_sample) first (the electron count comes from r_sample)")
        n = self._n_elec
        return self.nuclear_positions, np.ones(len(self.nuclear_positions)), (n + 1) // 2, n // 2, 0.0

    def _net_args(self):
        return dict(hidden=self.hidden_dims, envelope_decay=self.envelope_decay, activation=self.activation)

    def init(self, key, r_sample) -> Params:
        r_sample = np.asarray(r_sample)
        self._n_elec = int(r_sample.shape[0])
        rng = _generator(key)
        dims = [3 * self._n_elec, *self.hidden_dims, 1]                     # simple.py:63-68
        return {"params": {f"Dense_{i}": dense_params(rng, dims[i], dims[i + 1]) for i in range(len(dims) - 1)}}

    def sign_and_log(self, params, r, molecule=None):
        if self._n_elec is None:
            self._n_elec = int(np.asarray(r).shape[-2])
        return super().sign_and_log(params, r, molecule)

    @property
    def n_params(self) -> int:
        dims = [3 * (self._n_elec or 2), *self.hidden_dims, 1]              # simple.py:178-189
        return sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
