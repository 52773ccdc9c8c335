"""Shared helpers for the parity tests: seeded synthetic inputs in the shapes of SURVEY 8(d)."""
import numpy as np

from oracle.layout import KIND_ANTISYMMETRIC, KIND_SIMPLE, OracleNet, OracleSystem, hydrogen_chain_system, init_theta, param_count
from oracle.nqmc_oracle import Oracle
from oracle.nqmc_backflow import BackflowOracle


def make_oracle(s, net):
    return BackflowOracle(s, net) if net.use_backflow else Oracle(s, net)


def make_case(n_atoms=4, hidden=64, kind=KIND_ANTISYMMETRIC, use_cusp=False, W=512, seed=0, bond=1.8, burn=30,
              sigma=0.3, use_backflow=False):
    """-> (sys, net, theta(float32), r(float32, (W,N,3)) roughly |psi|^2-distributed)."""
    s = hydrogen_chain_system(n_atoms, bond)
    net = OracleNet(kind=kind, hidden=(hidden, hidden), use_cusp=use_cusp, use_backflow=use_backflow)
    theta = init_theta(s, net, seed + 1).astype(np.float32)
    rng = np.random.default_rng(seed)
    r = s.R[np.arange(s.n_elec) % s.n_atoms][None] + 0.5 * rng.standard_normal((W, s.n_elec, 3))   # vmc.py:199-208
    o = make_oracle(s, net)
    th64 = theta.astype(np.float64)
    _, la = o.log_psi(th64, r)
    logp = 2 * la
    for _ in range(burn):   # a few MH steps so walkers sit in |psi|^2 and away from nodes
        r, logp, _, _ = o.mh_step(th64, r, logp, rng.standard_normal(r.shape), rng.random(W), sigma)
    return s, net, theta, r.astype(np.float32)


def make_context(s, net):
    from nqmc_b200 import _native
    return _native.Context(s.R, s.Z, s.n_up, s.n_dn, s.v_nn, kind=net.kind, hidden=net.hidden, use_cusp=net.use_cusp,
                           use_backflow=net.use_backflow, envelope_decay=net.envelope_decay, activation=net.activation,
                           cusp_same=net.cusp_same, cusp_diff=net.cusp_diff)


def T(x):
    """DeviceArray -> torch CUDA tensor sharing its memory (tests only; the product never imports torch)."""
    import torch
    if isinstance(x, torch.Tensor):
        return x
    t = torch.as_tensor(x, device=torch.device("cuda", x.device))
    t._nqmc_keepalive = x
    return t


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(1.0, np.abs(b))
