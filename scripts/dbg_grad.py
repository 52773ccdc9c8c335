import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'neural-qmc-wavefunction_b200'))
import numpy as np, torch
from _util import make_case, make_context
from oracle.nqmc_oracle import Oracle
from oracle.layout import leaf_table, unflatten
for kw in (dict(n_atoms=4, hidden=64, W=64, seed=3), dict(n_atoms=6, hidden=64, W=384, use_cusp=True)):
    s, net, theta, r = make_case(**kw)
    ctx = make_context(s, net); ctx.set_params(theta)
    o = Oracle(s, net); th64 = theta.astype(np.float64); r64 = r.astype(np.float64)
    O = o.grad_log_psi(th64, r64)
    O32 = Oracle(s, net).grad_log_psi(theta, r)          # oracle in float32
    got = ctx.grad_accum(ctx.to_soa(r), None).cpu().numpy()
    ref = O.sum(0); scale = np.abs(O).sum(0)
    err = np.abs(got - ref) / np.maximum(scale, 1e-3)
    err32 = np.abs(O32.astype(np.float64).sum(0) - ref) / np.maximum(scale, 1e-3)
    tab = leaf_table(s, net)
    worst = np.argsort(err)[-5:]
    for i in worst:
        leaf = [p for p, sh, off in tab if off <= i][-1]
        wdom = np.argmax(np.abs(O[:, i]))
        print(kw, i, leaf, "err", err[i], "fp32-oracle err", err32[i], "got", got[i], "ref", ref[i], "dominant walker share", np.abs(O[wdom, i]) / scale[i])
    # conditioning
    sign, la, aux = o._forward(th64, r64, False)
    mats = aux[2]
    conds = np.max([np.linalg.cond(m[0]) for m in mats if m is not None], axis=0)
    print("cond quantiles", np.quantile(conds, [0.5, 0.9, 0.99, 1.0]))
    print("fp32 oracle max err", err32.max(), "device max err", err.max())
