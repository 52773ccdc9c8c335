"""Experiment: what a third / fourth A-staging buffer buys the producer/consumer forward kernel (NQMC_EXP modes of
nqmc_orbital_fwd_pc.cu: 1 = 4 channels x 2 buffers, 2 = 4 channels x 3 buffers, 4 = 4 channels x 4 buffers,
3 = 5 channels x 2 buffers (the generalised kernel at H = 64), 0 = production nqmc_orbital_lap_pc.cu)."""
import os, sys
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
import numpy as np
from oracle.layout import OracleNet, hydrogen_chain_system, init_theta
from _util import make_context
s = hydrogen_chain_system(4); net = OracleNet(hidden=(64, 64))
ctx = make_context(s, net); ctx.set_params(init_theta(s, net, 1).astype(np.float32))
W = 65536
r = (s.R[np.arange(4) % 4][None] + 0.5 * np.random.default_rng(0).standard_normal((W, 4, 3))).astype(np.float32)
soa = ctx.to_soa(r)
for mode in [0, 3, 1, 2, 4, 0]:
    os.environ["NQMC_EXP"] = str(mode)
    for _ in range(3): ctx.local_energy(soa)
    ctx.profile(True)
    for _ in range(10): ctx.local_energy(soa)
    ms, n = ctx.profile_read(1)
    ctx.profile(False)
    print(f"exp={mode}  eval5 {ms / n * 1e3:7.1f} us")
