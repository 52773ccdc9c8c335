"""Experiment: eval5 (orb_lap_pc_kernel) duration under debug modes (NQMC_DBG bit mask read per launch):
1 = hi*hi product only (40 instead of 120 UMMAs), 2 = no per-nucleus layer-1 FMAs, 4 = no split/STTM, 8 = no epilogue math."""
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
for mode in [0, 1, 2, 4, 8, 3, 6, 7, 14, 15]:
    os.environ["NQMC_DBG"] = str(mode)
    for _ in range(3): ctx.local_energy(soa)
    ctx.profile(True)
    for _ in range(10): ctx.local_energy(soa)
    ms, n = ctx.profile_read(1)
    ctx.profile(False)
    print(f"dbg={mode:2d}  eval5 {ms / n * 1e3:7.1f} us")
