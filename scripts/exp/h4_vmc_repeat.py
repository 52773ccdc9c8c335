import sys, numpy as np
sys.path[:0] = ["/root/repo", "/root/repo/tests", "/root/repo/neural-qmc-wavefunction_b200"]
import nqmc_b200 as nqmc
from nqmc_b200.optimizer import VMCOptimizer
from nqmc_b200.systems.molecule import hydrogen_chain
from nqmc_b200.wavefunction import AntisymmetricWavefunction
for rep in range(8):
    mol = hydrogen_chain(4, 1.8)
    wf = AntisymmetricWavefunction(2, 2, (64, 64), use_cusp=True, molecule=mol)
    params = wf.init(nqmc.random.PRNGKey(0))
    opt = VMCOptimizer(learning_rate=1e-3, n_samples=10, n_chains=1024, n_burn=100, step_size=0.5, clip_grad=1.0)
    params, hist = opt.train(nqmc.random.PRNGKey(1), wf, params, mol, n_steps=80, log_every=1000)
    e = np.array([h["energy"] for h in hist])
    print(rep, e[:5].mean(), e[-5:].mean(), e[:5].mean() - e[-5:].mean(), max(h["n_bad"] for h in hist))
