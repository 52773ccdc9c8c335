import sys, os
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path[:0] = [R, os.path.join(R, 'tests'), os.path.join(R, 'neural-qmc-wavefunction_b200')]
import numpy as np, torch
from _util import make_case, make_context, make_oracle, rel_err
s, net, theta, r = make_case(n_atoms=4, hidden=64, W=1000)
ctx = make_context(s, net); ctx.set_params(theta)
soa = ctx.to_soa(r)
o = make_oracle(s, net)
e_ref, parts, _, _ = o.local_energy(theta.astype(np.float64), r.astype(np.float64))
for tc in (0, 1):
    ctx.use_tensor_cores(bool(tc))
    e = ctx.local_energy(soa).cpu().numpy()
    torch.cuda.synchronize()
    err = rel_err(e, e_ref)
    print("tc", tc, "E_L rel err median/p99/max", np.median(err), np.quantile(err, .99), err.max(), "nan", np.isnan(e).sum())
