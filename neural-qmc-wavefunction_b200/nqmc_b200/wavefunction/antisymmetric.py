"""AntisymmetricWavefunction: det(Phi_up) det(Phi_dn) exp(J + cusps) (ansatz kind 1).

The reference only *specifies* this ansatz (.github/phase2-issue-body.md:78-200,
.github/phase3-issue-body.md:50-219, .github/phase4-issue-body.md:87-111).  The
constructor keywords follow README.md:185-190 / phase4-issue-body.md:87-90; the call contract is the
implemented ABC's (``wf(params, r) -> log|psi|``) with the sign available from ``sign_and_log``.

Parameter tree (what ``wf.init`` of the spec's flax modules would produce):
    params/orbitals_{up,down}/orbital_k/Dense_{0,1,2}/{kernel,bias}
    params/jastrow/{a_ee,b_ee,a_en,b_en}
    params/en_cusp/Dense_{2a,2a+1}/{kernel,bias}, params/ee_cusp/b_ee        (use_cusp)
    params/backflow/Dense_{0,1,2}/{kernel,bias}                                (use_backflow)
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np

from .. import _native
from ..random import _generator
from .base import Params, Wavefunction, dense_params

CUSP_HIDDEN = 16
BACKFLOW_HIDDEN = 16


class AntisymmetricWavefunction(Wavefunction):
    kind = _native.KIND_ANTISYMMETRIC

    def __init__(self, n_up: int, n_down: int, orbital_hidden_dims: Tuple[int, ...] = (32, 32), use_cusp: bool = False,
                 use_backflow: bool = False, molecule=None, nuclear_positions=None, charges=None,
                 cusp_same: float = 0.5, cusp_diff: float = 0.25, hidden_dims: Optional[Tuple[int, ...]] = None):
        if hidden_dims is not None:                       # phase2 spelling (phase2-issue-body.md:168)
            orbital_hidden_dims = hidden_dims
        if len(orbital_hidden_dims) != 2 or orbital_hidden_dims[0] != orbital_hidden_dims[1]:
            raise NotImplementedError("orbital_hidden_dims must be two equal widths (32, 64 or 128)")
        self.n_up, self.n_down = int(n_up), int(n_down)
        self.orbital_hidden_dims = tuple(int(h) for h in orbital_hidden_dims)
        self.use_cusp, self.use_backflow = bool(use_cusp), bool(use_backflow)
        self.cusp_same, self.cusp_diff = float(cusp_same), float(cusp_diff)
        self._v_nn = 0.0
        if molecule is not None:
            nuclear_positions, charges = molecule.positions, molecule.charges
            self._v_nn = float(molecule.nuclear_repulsion_energy())
            if tuple(molecule.spin) != (self.n_up, self.n_down):
                raise ValueError(f"molecule spin {molecule.spin} differs from (n_up, n_down)=({n_up},{n_down})")
        if nuclear_positions is None:
            raise ValueError("AntisymmetricWavefunction needs the geometry: pass molecule= or nuclear_positions=")
        self.nuclear_positions = np.asarray(nuclear_positions, dtype=np.float64).reshape(-1, 3)
        self._charges_given = charges is not None
        self.charges = (np.ones(len(self.nuclear_positions)) if charges is None
                        else np.asarray(charges, dtype=np.float64).reshape(-1))

    def _geometry(self, molecule=None):
        if molecule is not None:
            R = np.asarray(molecule.positions, dtype=np.float64)
            if R.shape != self.nuclear_positions.shape or not np.allclose(R, self.nuclear_positions):
                raise ValueError("molecule geometry differs from the wavefunction's nuclear_positions")
            Z = np.asarray(molecule.charges, dtype=np.float64).reshape(-1)
            if self._charges_given and not np.allclose(Z, self.charges):
                raise ValueError("molecule charges differ from the wavefunction's charges")
            # bind: from now on the sampler (which calls context() without a molecule) and the energy pass
            # (which passes one) resolve to the SAME native context -- same charges, same V_nn
            self.charges, self._charges_given = Z, True
            self._v_nn = float(molecule.nuclear_repulsion_energy())
            if tuple(molecule.spin) != (self.n_up, self.n_down):
                raise ValueError(f"molecule spin {molecule.spin} differs from (n_up, n_down)=({self.n_up},{self.n_down})")
        return self.nuclear_positions, self.charges, self.n_up, self.n_down, self._v_nn

    def _net_args(self):
        return dict(hidden=self.orbital_hidden_dims, use_cusp=self.use_cusp, use_backflow=self.use_backflow,
                    cusp_same=self.cusp_same, cusp_diff=self.cusp_diff)

    def init(self, key, r_sample=None) -> Params:
        rng = _generator(key)
        A = len(self.nuclear_positions)
        F = 3 + 2 * A
        h0, h1 = self.orbital_hidden_dims
        one = lambda: np.ones((), dtype=np.float32)

        def orbital():
            return {"Dense_0": dense_params(rng, F, h0), "Dense_1": dense_params(rng, h0, h1),
                    "Dense_2": dense_params(rng, h1, 1)}

        tree = {
            "orbitals_up": {f"orbital_{k}": orbital() for k in range(self.n_up)},
            "orbitals_down": {f"orbital_{k}": orbital() for k in range(self.n_down)},
            "jastrow": {"a_ee": one(), "b_ee": one(), "a_en": one(), "b_en": one()},     # phase2-issue-body.md:137-147
        }
        if self.use_cusp:
            en = {}
            for a in range(A):
                en[f"Dense_{2 * a}"] = dense_params(rng, 3, CUSP_HIDDEN)
                en[f"Dense_{2 * a + 1}"] = dense_params(rng, CUSP_HIDDEN, 1)
            tree["en_cusp"] = en
            tree["ee_cusp"] = {"b_ee": one()}                                             # phase3-issue-body.md:119
        if self.use_backflow:
            tree["backflow"] = {"Dense_0": dense_params(rng, 2, BACKFLOW_HIDDEN),
                                "Dense_1": dense_params(rng, BACKFLOW_HIDDEN, BACKFLOW_HIDDEN),
                                "Dense_2": dense_params(rng, BACKFLOW_HIDDEN, 1)}
        return {"params": tree}
