"""One fused walker step of the 128-wide configs under ncu's launch list (per-kernel durations)."""
import os, sys
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
import numpy as np
from oracle.layout import OracleNet, hydrogen_chain_system, init_theta
from _util import make_context
CASES = [("H6 h128 bf cusp", 6, 128, 32768, True, True), ("H10 h128", 10, 128, 131072, False, False)]
for name, n_atoms, hidden, W, cusp, bf in CASES:
    s = hydrogen_chain_system(n_atoms); net = OracleNet(hidden=(hidden, hidden), use_cusp=cusp, use_backflow=bf)
    ctx = make_context(s, net); ctx.set_params(init_theta(s, net, 1).astype(np.float32))
    r = (s.R[np.arange(s.n_elec) % s.n_atoms][None] + 0.5 * np.random.default_rng(0).standard_normal((W, s.n_elec, 3))).astype(np.float32)
    soa = ctx.to_soa(r); _, la = ctx.log_psi(soa)
    logp = ctx.upload((2 * la.numpy()).astype(np.float32))
    e, h, g = ctx.empty(W), ctx.empty(8, np.float64), ctx.empty(ctx.P, np.float64)
    for t in range(3): ctx.walker_step(soa, logp, 0.3, e, h, g, seed=3, step=t)
    ctx.synchronize()
    del ctx
