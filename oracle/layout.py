"""ORACLE (test infrastructure only) -- parameter trees and the flat theta vector.

This file is part of the CPU oracle.  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import it.  The product
(neural-qmc-wavefunction_b200/) never does.

PARITY UNPINNED for the Slater/Jastrow/cusp/backflow ansatz: the reference has
no tests, fixtures or golden vectors for it, and JAX/Flax are not installed, so
nothing here could be checked against a run of the reference itself.  What is
pinned: geometry goldens from the reference's own molecule.py (tests/golden/).

What it restates
----------------
* flax ``nn.Dense`` parameter naming / shapes used by the reference
  (``src/nqmc/wavefunction/simple.py:41-70``: ``Dense_i`` with ``kernel (in,out)``
  and ``bias (out,)``; ``y = x @ kernel + bias``).
* the [SPEC] module trees of ``.github/phase2-issue-body.md:78-113,125-152,164-176``
  (``orbitals_up/orbital_k``, ``jastrow/{a_ee,b_ee,a_en,b_en}``) and
  ``.github/phase3-issue-body.md:50-83,95-122,134-169,180-195``
  (``en_cusp``, ``ee_cusp``, ``backflow``).
* the flat order: ``jax.tree.leaves`` walks dict keys in sorted order
  (used by the reference at ``src/nqmc/optimizer/vmc.py:265``), so theta is the
  concatenation of row-major leaves in sorted-key order -- exactly what
  ``jax.flatten_util.ravel_pytree(params)`` would give.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

KIND_SIMPLE = 0          # SimpleWavefunction (simple.py:73)
KIND_ANTISYMMETRIC = 1   # [SPEC] AntisymmetricWavefunction (phase3-issue-body.md:180)

BF_HIDDEN = 16           # phase3-issue-body.md:139  hidden_dims=(16,16)
CUSP_HIDDEN = 16         # phase3-issue-body.md:76   nn.Dense(16)


@dataclass
class OracleSystem:
    """Geometry constants the path needs (systems/molecule.py:49-97,142-153)."""
    R: np.ndarray            # (A,3) float64, bohr
    Z: np.ndarray            # (A,)  float64
    n_up: int
    n_dn: int
    v_nn: float

    @property
    def n_atoms(self) -> int:
        return int(self.R.shape[0])

    @property
    def n_elec(self) -> int:
        return self.n_up + self.n_dn


@dataclass
class OracleNet:
    kind: int = KIND_ANTISYMMETRIC
    hidden: Tuple[int, ...] = (64, 64)   # kind 0: any number of layers (simple.py:63-65); kind 1: two equal widths
    activation: str = "tanh"             # kind 0 only: 'tanh' or 'silu' (simple.py:56-58)
    use_cusp: bool = False
    use_backflow: bool = False
    envelope_decay: float = 1.0      # kind 0 only (simple.py:102)
    cusp_same: float = 0.5           # phase3-issue-body.md:109 (spec states these reversed vs Kato)
    cusp_diff: float = 0.25          # phase3-issue-body.md:110


def hydrogen_chain_system(n_atoms: int, bond: float = 1.8) -> OracleSystem:
    """systems/molecule.py:201-225 + :142-153 restated (nuclei at (i*R,0,0), Z=1)."""
    if n_atoms % 2 != 0 and n_atoms != 1:
        raise ValueError("n_atoms must be even")
    R = np.array([[i * bond, 0.0, 0.0] for i in range(n_atoms)], dtype=np.float64)
    Z = np.ones(n_atoms, dtype=np.float64)
    v = 0.0
    for i in range(n_atoms):
        for j in range(i + 1, n_atoms):
            v += Z[i] * Z[j] / np.linalg.norm(R[i] - R[j])
    n_up = 1 if n_atoms == 1 else n_atoms // 2     # molecule.py:83-84 (singlet; H atom: doublet)
    n_dn = n_atoms - n_up
    return OracleSystem(R=R, Z=Z, n_up=n_up, n_dn=n_dn, v_nn=float(v))


def _dense(n_in: int, n_out: int) -> Dict[str, Tuple[int, ...]]:
    return {"bias": (n_out,), "kernel": (n_in, n_out)}


def param_shapes(sysm: OracleSystem, net: OracleNet) -> Dict:
    """Nested dict of leaf shapes, mirroring the flax trees described above."""
    if net.kind == KIND_SIMPLE:
        dims = [3 * sysm.n_elec, *net.hidden, 1]                     # simple.py:63-68
        return {"params": {f"Dense_{i}": _dense(dims[i], dims[i + 1]) for i in range(len(dims) - 1)}}
    h0, h1 = net.hidden
    F = 3 + 2 * sysm.n_atoms
    mlp = lambda: {"Dense_0": _dense(F, h0), "Dense_1": _dense(h0, h1), "Dense_2": _dense(h1, 1)}
    tree: Dict = {
        "jastrow": {"a_ee": (), "a_en": (), "b_ee": (), "b_en": ()},
        "orbitals_up": {f"orbital_{k}": mlp() for k in range(sysm.n_up)},
        "orbitals_down": {f"orbital_{k}": mlp() for k in range(sysm.n_dn)},
    }
    if net.use_cusp:
        en = {}
        for a in range(sysm.n_atoms):      # flax auto-names inside the python loop
            en[f"Dense_{2 * a}"] = _dense(3, CUSP_HIDDEN)
            en[f"Dense_{2 * a + 1}"] = _dense(CUSP_HIDDEN, 1)
        tree["en_cusp"] = en
        tree["ee_cusp"] = {"b_ee": ()}
    if net.use_backflow:
        tree["backflow"] = {"Dense_0": _dense(2, BF_HIDDEN), "Dense_1": _dense(BF_HIDDEN, BF_HIDDEN),
                            "Dense_2": _dense(BF_HIDDEN, 1)}
    return {"params": tree}


def _walk(tree, prefix=""):
    for key in sorted(tree.keys()):
        v = tree[key]
        path = f"{prefix}/{key}" if prefix else key
        if isinstance(v, dict):
            yield from _walk(v, path)
        else:
            yield path, v


def leaf_table(sysm: OracleSystem, net: OracleNet) -> List[Tuple[str, Tuple[int, ...], int]]:
    """[(path, shape, offset)] in sorted-key (jax.tree.leaves) order."""
    out, off = [], 0
    for path, shape in _walk(param_shapes(sysm, net)):
        out.append((path, tuple(shape), off))
        off += int(np.prod(shape)) if len(shape) else 1
    return out


def param_count(sysm: OracleSystem, net: OracleNet) -> int:
    t = leaf_table(sysm, net)
    p, s, o = t[-1]
    return o + (int(np.prod(s)) if len(s) else 1)


def unflatten(sysm: OracleSystem, net: OracleNet, theta: np.ndarray) -> Dict[str, np.ndarray]:
    """flat theta -> {path: array} (views)."""
    out = {}
    for path, shape, off in leaf_table(sysm, net):
        n = int(np.prod(shape)) if len(shape) else 1
        out[path] = theta[off:off + n].reshape(shape)
    return out


def init_theta(sysm: OracleSystem, net: OracleNet, seed: int = 0, dtype=np.float64) -> np.ndarray:
    """LeCun-normal kernels (std 1/sqrt(fan_in)), zero biases (flax Dense defaults),
    Jastrow / cusp scalars = 1 (phase2-issue-body.md:137-147, phase3-issue-body.md:119)."""
    rng = np.random.default_rng(seed)
    theta = np.zeros(param_count(sysm, net), dtype=np.float64)
    for path, shape, off in leaf_table(sysm, net):
        n = int(np.prod(shape)) if len(shape) else 1
        leaf = path.rsplit("/", 1)[-1]
        if leaf == "kernel":
            # flax default: lecun_normal = truncated normal (+-2 sigma) scaled to unit variance
            std = 1.0 / np.sqrt(shape[0])
            x = rng.standard_normal(n * 2)
            x = x[np.abs(x) <= 2.0][:n]
            theta[off:off + n] = x * std / 0.87962566103423978
        elif leaf == "bias":
            theta[off:off + n] = 0.0
        else:
            theta[off:off + n] = 1.0
    return theta.astype(dtype)
