"""Throughput of the BASELINE configs that are parity-test cases (not the bench line), at their per-GPU sizes:
one fused walker step (MH + E_L + gradient sums), device time, L2 not flushed.  Prints walker-steps/s per GPU."""
import os
import sys

R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
import numpy as np  # noqa: E402
import torch  # noqa: E402
from oracle.layout import OracleNet, hydrogen_chain_system, init_theta  # noqa: E402
from _util import T, make_context  # noqa: E402

CASES = [("configs[1] H4 64-wide", 4, 64, 65536, False, False),
         ("configs[2] H6 +backflow +cusp, 128-wide", 6, 128, 32768, True, True),
         ("H6 64-wide (tensor-core path, A = 6)", 6, 64, 32768, False, False),
         ("configs[4] H10 128-wide", 10, 128, 131072, False, False)]
for name, n_atoms, hidden, W, cusp, bf in CASES:
    s = hydrogen_chain_system(n_atoms)
    net = OracleNet(hidden=(hidden, hidden), use_cusp=cusp, use_backflow=bf)
    theta = init_theta(s, net, 1).astype(np.float32)
    ctx = make_context(s, net)
    ctx.set_params(theta)
    rng = np.random.default_rng(0)
    r = (s.R[np.arange(s.n_elec) % s.n_atoms][None] + 0.5 * rng.standard_normal((W, s.n_elec, 3))).astype(np.float32)
    soa = ctx.to_soa(r)
    _, la = ctx.log_psi(soa)
    logp = (2 * T(la)).contiguous()
    dev = ctx.dev()
    e_l = torch.empty(W, dtype=torch.float32, device=dev)
    hdr = torch.empty(8, dtype=torch.float64, device=dev)
    gsum = torch.empty(ctx.P, dtype=torch.float64, device=dev)
    for t in range(5):
        ctx.walker_step(soa, logp, 0.3, e_l, hdr, gsum, seed=3, step=t)
    torch.cuda.synchronize()
    n = 10
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for t in range(n):
        ctx.walker_step(soa, logp, 0.3, e_l, hdr, gsum, seed=3, step=100 + t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:45s} W={W:7d} P={ctx.P:7d}  {ms:8.3f} ms/step  {W / ms * 1e3 / 1e6:8.2f} M walker-steps/s")
    del ctx
