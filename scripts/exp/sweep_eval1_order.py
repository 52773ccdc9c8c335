"""Experiment: stage order of the value kernel's three work items per quarter (NQMC_EVAL1_ORDER, 2 bits per quarter:
0 = a, b, c; 1 = b, a, c; 2 = b, c, a).  One subprocess per setting (the variable is read once per process)."""
import os, subprocess, sys
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
CHILD = r'''
import os, sys
R = %r
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
import numpy as np
from oracle.layout import OracleNet, hydrogen_chain_system, init_theta
from _util import make_context
s = hydrogen_chain_system(4); net = OracleNet(hidden=(64, 64))
ctx = make_context(s, net); ctx.set_params(init_theta(s, net, 1).astype(np.float32))
W = 65536
r = (s.R[np.arange(4) %% 4][None] + 0.5 * np.random.default_rng(0).standard_normal((W, 4, 3))).astype(np.float32)
soa = ctx.to_soa(r)
for _ in range(5): ctx.log_psi(soa)
ctx.profile(True)
for _ in range(20): ctx.log_psi(soa)
ms, n = ctx.profile_read(0)
print("order", os.environ.get("NQMC_EVAL1_ORDER"), "eval1 %%.1f us" %% (ms / n * 1e3))
''' % R
masks = sys.argv[1:] or ["0x00", "0x44", "0x11", "0x50", "0x05", "0x64", "0x46", "0x98", "0x89", "0x24", "0x18", "0xa0", "0x60", "0x90"]
for m in masks:
    env = dict(os.environ, NQMC_EVAL1_ORDER=m)
    print(subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True).stdout.strip())
