"""Molecular geometry for the walker loop (host side, NumPy only).

Same public surface as the reference's ``nqmc.systems.molecule`` (src/nqmc/systems/molecule.py:
``Molecule`` :27-182, ``hydrogen_molecule`` :185, ``hydrogen_chain`` :201, ``hydrogen_atom`` :228).
This layer is not accelerated: it only supplies constants (R_A, Z_A, spin split, V_nn) to the
device context.  tests/test_molecule_golden.py checks it against values produced by executing the
reference's own file (tests/golden/molecule_golden.json).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence, Tuple

import numpy as np

ELEMENT_CHARGES = {"H": 1, "He": 2, "Li": 3, "Be": 4, "B": 5, "C": 6, "N": 7, "O": 8}
BOHR_TO_ANGSTROM = 0.529177


@dataclass
class Molecule:
    atoms: List[Tuple[str, np.ndarray]]
    charges: np.ndarray
    n_electrons: int
    spin: Tuple[int, int]

    @classmethod
    def from_atoms(cls, elements: Sequence[str], positions, charge: int = 0, spin_polarization: int = 0) -> "Molecule":
        if len(elements) != len(positions):
            raise ValueError(f"Length mismatch: {len(elements)} elements vs {len(positions)} positions")
        z = np.array([ELEMENT_CHARGES[e] for e in elements], dtype=np.float64)     # KeyError for unknown symbols
        xyz = [np.array(p, dtype=np.float64) for p in positions]
        n_el = int(z.sum()) - charge
        n_up = (n_el + spin_polarization) // 2
        n_dn = n_el - n_up
        if n_up < 0 or n_dn < 0:
            raise ValueError(f"Invalid spin configuration: n_up={n_up}, n_down={n_dn}")
        return cls(atoms=list(zip(elements, xyz)), charges=z, n_electrons=n_el, spin=(n_up, n_dn))

    @property
    def n_atoms(self) -> int:
        return len(self.atoms)

    @property
    def positions(self) -> np.ndarray:
        return np.array([p for _, p in self.atoms])

    @property
    def elements(self) -> List[str]:
        return [e for e, _ in self.atoms]

    def electron_nuclear_distances(self, electron_positions: np.ndarray) -> np.ndarray:
        d = np.asarray(electron_positions)[:, None, :] - self.positions[None, :, :]
        return np.sqrt((d * d).sum(-1))

    def electron_electron_distances(self, electron_positions: np.ndarray) -> np.ndarray:
        r = np.asarray(electron_positions)
        d = r[:, None, :] - r[None, :, :]
        return np.sqrt((d * d).sum(-1))

    def nuclear_repulsion_energy(self) -> float:
        R, Z = self.positions, self.charges
        e = 0.0
        for a in range(self.n_atoms):
            for b in range(a + 1, self.n_atoms):
                e += Z[a] * Z[b] / np.linalg.norm(R[a] - R[b])
        return e

    def to_xyz_string(self) -> str:
        rows = [str(self.n_atoms), ""]
        for el, p in self.atoms:
            q = p * BOHR_TO_ANGSTROM
            rows.append(f"{el} {q[0]:.6f} {q[1]:.6f} {q[2]:.6f}")
        return "\n".join(rows)

    def get_spin_mask(self) -> np.ndarray:
        n_up, _ = self.spin
        up = np.arange(self.n_electrons) < n_up
        return up[:, None] == up[None, :]


def hydrogen_molecule(bond_length: float = 1.4) -> Molecule:
    return Molecule.from_atoms(["H", "H"], np.array([[0.0, 0.0, 0.0], [bond_length, 0.0, 0.0]]))


def hydrogen_chain(n_atoms: int, bond_length: float = 1.8) -> Molecule:
    if n_atoms % 2 != 0:
        raise ValueError(f"n_atoms must be even for closed-shell hydrogen chain, got {n_atoms}")
    xyz = np.zeros((n_atoms, 3))
    xyz[:, 0] = np.arange(n_atoms) * bond_length
    return Molecule.from_atoms(["H"] * n_atoms, xyz)


def hydrogen_atom() -> Molecule:
    return Molecule.from_atoms(["H"], np.zeros((1, 3)), spin_polarization=1)
