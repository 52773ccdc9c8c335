"""SimpleWavefunction: MLP(flatten r) + exponential envelope (ansatz kind 0).

Host-side mirror of src/nqmc/wavefunction/simple.py:73-176 -- same constructor, same parameter tree
(``params/Dense_{0,1,2}/{kernel,bias}``) -- evaluated by csrc/nqmc_simple.cu.  Only 'tanh' is
implemented natively; 'silu' (simple.py:58) raises.
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
        if activation != "tanh":
            raise NotImplementedError("the native SimpleWavefunction kernel implements activation='tanh' only")
        if len(hidden_dims) != 2 or hidden_dims[0] != hidden_dims[1]:
            raise NotImplementedError("the native SimpleWavefunction kernel needs two equal hidden widths")
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
            return R, np.asarray(molecule.charges, dtype=np.float64), int(n_up), int(n_dn), float(molecule.nuclear_repulsion_energy())
        if self._n_elec is None:
            raise RuntimeError("call init(key, r_sample) first (the electron count comes from r_sample)")
        n = self._n_elec
        return self.nuclear_positions, np.ones(len(self.nuclear_positions)), (n + 1) // 2, n // 2, 0.0

    def _net_args(self):
        return dict(hidden=self.hidden_dims, envelope_decay=self.envelope_decay)

    def init(self, key, r_sample) -> Params:
        r_sample = np.asarray(r_sample)
        self._n_elec = int(r_sample.shape[0])
        rng = _generator(key)
        d = 3 * self._n_elec
        h0, h1 = self.hidden_dims
        return {"params": {"Dense_0": dense_params(rng, d, h0), "Dense_1": dense_params(rng, h0, h1),
                           "Dense_2": dense_params(rng, h1, 1)}}

    def sign_and_log(self, params, r, molecule=None):
        if self._n_elec is None:
            self._n_elec = int(np.asarray(r).shape[-2])
        return super().sign_and_log(params, r, molecule)

    @property
    def n_params(self) -> int:
        n_in = 3 * (self._n_elec or 2)
        h0, h1 = self.hidden_dims
        return n_in * h0 + h0 + h0 * h1 + h1 + h1 + 1
