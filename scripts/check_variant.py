"""Parity of whichever kernel variants the environment selects (NQMC_LAP_PC, NQMC_LAP_TC, NQMC_TC_BWD, NQMC_TC are
read once per process): E_L and the weighted gradient sum on 1000 H4 walkers against the fp64 oracle.
Run by tests/test_gpu_variants.py in a subprocess per variant; prints VARIANT_OK on success."""
import os
import sys

R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
import numpy as np  # noqa: E402
from _util import make_case, make_context, make_oracle, rel_err  # noqa: E402

s, net, theta, r = make_case(n_atoms=4, hidden=64, W=1000)
ctx = make_context(s, net)
ctx.set_params(theta)
soa = ctx.to_soa(r)
o = make_oracle(s, net)
th64, r64 = theta.astype(np.float64), r.astype(np.float64)
e_ref, _, _, la_ref = o.local_energy(th64, r64)
e = ctx.local_energy(soa).cpu().numpy()
_, la = ctx.log_psi(soa)
err_e = rel_err(e, e_ref)
err_l = rel_err(la.cpu().numpy(), la_ref)
O = o.grad_log_psi(th64, r64)
wgt = np.random.default_rng(5).standard_normal(1000).astype(np.float32)
ref = (wgt.astype(np.float64)[:, None] * O).sum(0)
scale = (np.abs(wgt)[:, None] * np.abs(O)).sum(0)
got = ctx.grad_accum(soa, ctx.upload(wgt)).cpu().numpy()
err_g = np.abs(got - ref) / np.maximum(scale, 1e-3)
print(f"E_L rel err median {np.median(err_e):.2e} max {err_e.max():.2e}; log|psi| max {err_l.max():.2e}; "
      f"grad median {np.median(err_g):.2e} p99 {np.quantile(err_g, .99):.2e} max {err_g.max():.2e}")
ok = (np.isfinite(e).all() and np.median(err_e) < 2e-6 and err_e.max() < 1e-4 and err_l.max() < 1e-5
      and np.median(err_g) < 1e-6 and np.quantile(err_g, .99) < 1e-5 and err_g.max() < 1e-3)
print("VARIANT_OK" if ok else "VARIANT_FAIL")
sys.exit(0 if ok else 1)
