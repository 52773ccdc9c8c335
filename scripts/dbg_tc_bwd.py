import sys, os
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path[:0] = [R, os.path.join(R, 'tests'), os.path.join(R, 'neural-qmc-wavefunction_b200')]
import numpy as np, torch
from _util import make_case, make_context, make_oracle, rel_err
from oracle.layout import leaf_table
s, net, theta, r = make_case(n_atoms=4, hidden=64, W=1000)
ctx = make_context(s, net); ctx.set_params(theta)
soa = ctx.to_soa(r)
o = make_oracle(s, net)
O = o.grad_log_psi(theta.astype(np.float64), r.astype(np.float64))
rng = np.random.default_rng(5)
wgt = rng.standard_normal(1000).astype(np.float32)
ref = (wgt.astype(np.float64)[:, None] * O).sum(0); scale = (np.abs(wgt)[:, None] * np.abs(O)).sum(0)
tab = leaf_table(s, net)
for tc in (0, 1):
    ctx.use_tensor_cores(bool(tc))
    got = ctx.grad_accum(soa, torch.as_tensor(wgt).cuda()).cpu().numpy()
    torch.cuda.synchronize()
    err = np.abs(got - ref) / np.maximum(scale, 1e-3)
    print("tc", tc, "grad err median/p99/max", np.median(err), np.quantile(err, .99), err.max(), "nan", np.isnan(got).sum())
    for path, sh, off in tab:
        n = int(np.prod(sh)) if len(sh) else 1
        e = err[off:off+n]
        if e.max() > 1e-3: print("   bad leaf", path, e.max(), got[off:off+3], ref[off:off+3])
