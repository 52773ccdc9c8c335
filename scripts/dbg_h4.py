import sys, os
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path[:0] = [R, os.path.join(R, 'neural-qmc-wavefunction_b200')]
import numpy as np
import nqmc_b200 as nqmc
from nqmc_b200.optimizer import VMCOptimizer
from nqmc_b200.systems.molecule import hydrogen_chain
from nqmc_b200.wavefunction import AntisymmetricWavefunction
mol = hydrogen_chain(4, 1.8)
for lr in (1e-3, 1e-2):
    wf = AntisymmetricWavefunction(2, 2, (64, 64), use_cusp=True, molecule=mol)
    params = wf.init(nqmc.random.PRNGKey(0))
    opt = VMCOptimizer(learning_rate=lr, n_samples=10, n_chains=1024, n_burn=100, step_size=0.5, clip_grad=1.0)
    params, hist = opt.train(nqmc.random.PRNGKey(1), wf, params, mol, n_steps=150, log_every=1000)
    e = np.array([h["energy"] for h in hist]); s = np.array([h["energy_std"] for h in hist]); g=np.array([h["grad_norm"] for h in hist]); a=np.array([h["acceptance_rate"] for h in hist])
    print("lr", lr, "E:", np.round(e[::10], 3), "std", np.round(s[::10], 3), "gn", np.round(g[::10],2), "acc", np.round(a[::20],2))
