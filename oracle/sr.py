"""Oracle for Stochastic Reconfiguration -- TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

NumPy fp64 restatement of the reference's [SPEC] StochasticReconfiguration.compute_update,
/root/reference/.github/phase6-issue-body.md:150-181 (the reference ships no src/nqmc/optimizer/sr.py; the
issue body's code block is the only definition).  Pinned by tests/golden/ref_sr_golden.npz, which
tests/golden/make_ref_sr_golden.py produced by EXECUTING that code block under the jax stand-in.
"""
import numpy as np


def sr_matrices(log_psi_grads, local_energies, reg):
    """phase6-issue-body.md:163-172: centred O and E, S = O^T O / n + reg I, f = 2 E^T O / n."""
    g = np.asarray(log_psi_grads, dtype=np.float64)
    e = np.asarray(local_energies, dtype=np.float64)
    n, p = g.shape
    O = g - g.mean(axis=0)                      # :164
    E = e - e.mean()                            # :165
    S = np.einsum("ni,nj->ij", O, O) / n        # :168
    S = S + reg * np.eye(p)                     # :169
    f = 2.0 * np.einsum("n,ni->i", E, O) / n    # :172
    return S, f


def sr_update(log_psi_grads, local_energies, learning_rate=1e-2, regularization=1e-4):
    """phase6-issue-body.md:150-181: -lr * solve(S, f)."""
    S, f = sr_matrices(log_psi_grads, local_energies, regularization)
    return -learning_rate * np.linalg.solve(S, f)   # :175-177
