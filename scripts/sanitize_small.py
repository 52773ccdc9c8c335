"""Small end-to-end exercise of every kernel family for compute-sanitizer (memcheck): tiny walker counts."""
import sys, os
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path[:0] = [R, os.path.join(R, 'tests'), os.path.join(R, 'neural-qmc-wavefunction_b200')]
import numpy as np, torch
from _util import make_case, make_context
from oracle.layout import KIND_SIMPLE
cases = [dict(n_atoms=4, hidden=64, W=200), dict(n_atoms=6, hidden=64, W=130, use_cusp=True, use_backflow=True),
         dict(n_atoms=2, hidden=32, W=70), dict(n_atoms=4, hidden=128, W=90), dict(n_atoms=2, hidden=64, W=50, kind=KIND_SIMPLE)]
for kw in cases:
    s, net, theta, r = make_case(burn=2, **kw)
    for tc in (1, 0):
        ctx = make_context(s, net); ctx.use_tensor_cores(bool(tc)); ctx.set_params(theta)
        W = r.shape[0]; dev = ctx.dev()
        soa = ctx.to_soa(r)
        sg, la = ctx.log_psi(soa)
        logp = (2 * la).contiguous()
        e_l = torch.empty(W, dtype=torch.float32, device=dev); hdr = torch.empty(8, dtype=torch.float64, device=dev)
        gsum = torch.empty(ctx.P, dtype=torch.float64, device=dev)
        ctx.walker_step(soa, logp, 0.3, e_l, hdr, gsum, seed=1, step=2)
        ctx.grad_accum(soa, None)
        ctx.local_energy(soa, want_parts=True)
        torch.cuda.synchronize()
        print(kw, "tc", tc, "ok", float(hdr[1] / hdr[0]))
