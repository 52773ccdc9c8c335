#!/usr/bin/env python
"""H2 VMC with the reference's script defaults (scripts/train_h2.py:39-78,165-201): SimpleWavefunction (64,64),
32 chains x 300 samples, 300 burn-in, step 0.2, lr 1e-2, clip 1.0, 200 iterations, seed 42.
Exit code 0 when the mean of the last <= 50 energies is below -1.10 Ha (scripts/train_h2.py:206-216).
No plotting (matplotlib is out of scope); prints the energy trace instead."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "neural-qmc-wavefunction_b200"))

import nqmc_b200 as nqmc
from nqmc_b200.optimizer import VMCOptimizer
from nqmc_b200.systems.molecule import hydrogen_molecule
from nqmc_b200.wavefunction import AntisymmetricWavefunction, SimpleWavefunction

EXACT_H2 = -1.174


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n_steps", type=int, default=200)
    ap.add_argument("--lr", type=float, default=1e-2)
    ap.add_argument("--hidden_dims", default="64,64")
    ap.add_argument("--n_samples", type=int, default=300)
    ap.add_argument("--n_chains", type=int, default=32)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--bond_length", type=float, default=1.4)
    ap.add_argument("--log_every", type=int, default=10)
    ap.add_argument("--ansatz", default="simple", choices=["simple", "slater"])
    ap.add_argument("--optimizer", default="adam", choices=["adam", "sr"],
                    help="sr: Stochastic Reconfiguration on the device (simple ansatz; try --lr 0.05 --n_chains 512 --n_samples 8)")
    ap.add_argument("--sr_regularization", type=float, default=1e-3)
    a = ap.parse_args()
    hidden = tuple(int(x) for x in a.hidden_dims.split(","))
    mol = hydrogen_molecule(a.bond_length)
    key = nqmc.random.PRNGKey(a.seed)
    key, init_key, train_key = nqmc.random.split(key, 3)
    if a.ansatz == "simple":
        wf = SimpleWavefunction(hidden_dims=hidden, envelope_decay=1.0, nuclear_positions=mol.positions)
        params = wf.init(init_key, np.zeros((2, 3)))
    else:
        wf = AntisymmetricWavefunction(1, 1, hidden, molecule=mol)
        params = wf.init(init_key)
    opt = VMCOptimizer(learning_rate=a.lr, n_samples=a.n_samples, n_chains=a.n_chains, n_burn=300, step_size=0.2, clip_grad=1.0,
                       optimizer=a.optimizer, sr_regularization=a.sr_regularization)
    t0 = time.time()
    params, history = opt.train(train_key, wf, params, mol, n_steps=a.n_steps, log_every=a.log_every)
    dt = time.time() - t0
    e = np.array([h["energy"] for h in history])
    final = float(np.mean(e[-min(50, len(e)):]))
    print(f"final <E> over last {min(50, len(e))} iterations: {final:.5f} Ha (exact {EXACT_H2}); "
          f"{a.n_steps} iterations in {dt:.1f} s; acceptance {history[-1]['acceptance_rate']:.2f}")
    return 0 if final < -1.10 else 1


if __name__ == "__main__":
    sys.exit(main())
