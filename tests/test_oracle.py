"""CPU tests of the oracle itself: the two independent restatements agree, and the known-answer
tests the reference's documents sanction hold."""
import numpy as np
import pytest
import torch

from oracle.layout import (KIND_ANTISYMMETRIC, KIND_SIMPLE, OracleNet, hydrogen_chain_system, init_theta, leaf_table,
                           param_count, unflatten)
from oracle.nqmc_autodiff import AutodiffOracle
from oracle.nqmc_oracle import Oracle
from oracle import philox


def _case(n_atoms, hidden=16, kind=KIND_ANTISYMMETRIC, cusp=False, W=12, seed=0, bf=False):
    s = hydrogen_chain_system(n_atoms)
    net = OracleNet(kind=kind, hidden=(hidden, hidden), use_cusp=cusp, use_backflow=bf)
    th = init_theta(s, net, seed + 1) + 0.05 * np.random.default_rng(seed + 2).standard_normal(param_count(s, net))
    r = s.R[np.arange(s.n_elec) % s.n_atoms][None] + 0.5 * np.random.default_rng(seed).standard_normal((W, s.n_elec, 3))
    return s, net, th, r


@pytest.mark.parametrize("n_atoms,kind,cusp", [(1, 1, False), (2, 1, False), (4, 1, False), (4, 1, True), (6, 1, True),
                                               (2, 0, False), (4, 0, False)])
def test_analytic_equals_autodiff_fp64(n_atoms, kind, cusp):
    """K-8: forward-Laplacian formulas (Appendix C) == nested autodiff, to 1e-9 in float64."""
    s, net, th, r = _case(n_atoms, kind=kind, cusp=cusp)
    o, a = Oracle(s, net), AutodiffOracle(s, net)
    sg, la = o.log_psi(th, r)
    sg2, la2 = a.log_psi(th, r)
    assert np.array_equal(sg, sg2)
    assert np.max(np.abs(la - la2)) < 1e-10
    e, _, _, _ = o.local_energy(th, r)
    e2 = a.local_energy(th, r)
    assert np.max(np.abs(e - e2) / np.maximum(1, np.abs(e2))) < 1e-9
    O, O2 = o.grad_log_psi(th, r), a.grad_log_psi(th, r)
    assert np.max(np.abs(O - O2) / np.maximum(1, np.abs(O2))) < 1e-9


def test_param_counts_match_survey_table():
    """P = 19,976 (H4/64), 112,962 (H6/128 +BF+cusp), 197,134 (H10/128)."""
    assert param_count(hydrogen_chain_system(4), OracleNet(hidden=(64, 64))) == 19976
    assert param_count(hydrogen_chain_system(6), OracleNet(hidden=(128, 128), use_cusp=True, use_backflow=True)) == 112962
    assert param_count(hydrogen_chain_system(10), OracleNet(hidden=(128, 128))) == 197134
    assert param_count(hydrogen_chain_system(2), OracleNet(kind=KIND_SIMPLE, hidden=(64, 64))) == 4673   # simple.py:178-189


def test_h_atom_exact_local_energy():
    """K-1 (docs/IMPLEMENTATION_PLAN.md:350-364): psi = exp(-r) => E_L == -0.5 pointwise.
    SimpleWavefunction with a zero MLP is exactly exp(-|r|) for one nucleus at the origin, alpha = 1."""
    s = hydrogen_chain_system(1)
    net = OracleNet(kind=KIND_SIMPLE, hidden=(8, 8))
    th = np.zeros(param_count(s, net))
    r = np.random.default_rng(0).standard_normal((64, 1, 3))
    e, parts, _, la = Oracle(s, net).local_energy(th, r)
    assert np.max(np.abs(e + 0.5)) < 1e-12
    assert np.allclose(la, -np.linalg.norm(r[:, 0], axis=-1))
    e2 = AutodiffOracle(s, net).local_energy(th, r)
    assert np.max(np.abs(e2 + 0.5)) < 1e-10


def test_metropolis_gaussian_variance():
    """K-2 (IMPLEMENTATION_PLAN.md:366-382): MH on log p = -2 sum r^2 => per-coordinate variance 0.25 +- 0.05."""
    s, net, th, _ = _case(2)
    o = Oracle(s, net)
    o.log_psi = lambda theta, r: (np.ones(r.shape[0]), -np.sum(r * r, axis=(-1, -2)))     # log_prob = 2 log|psi|
    rng = np.random.default_rng(0)
    r = rng.standard_normal((256, 2, 3)) * 0.5
    logp = 2 * o.log_psi(th, r)[1]
    acc, samples = [], []
    for t in range(400):
        r, logp, a, _ = o.mh_step(th, r, logp, rng.standard_normal(r.shape), rng.random(256), 0.4)
        acc.append(a.mean())
        if t >= 100:
            samples.append(r.copy())
    var = np.var(np.stack(samples))
    assert abs(var - 0.25) < 0.05, var
    assert 0.2 < np.mean(acc) < 0.9


@pytest.mark.parametrize("n_atoms,cusp", [(4, False), (6, True)])
def test_antisymmetry(n_atoms, cusp):
    """K-3 (phase2-issue-body.md:212-230): same-spin swap => log|psi| equal (rtol 1e-5), sign flips."""
    s, net, th, r = _case(n_atoms, cusp=cusp)
    o = Oracle(s, net)
    sg, la = o.log_psi(th, r)
    sw = r.copy(); sw[:, [0, 1]] = sw[:, [1, 0]]                     # both spin-up
    sg2, la2 = o.log_psi(th, sw)
    assert np.allclose(la, la2, rtol=1e-5)
    assert np.array_equal(sg, -sg2)
    dn = r.copy(); dn[:, [s.n_up, s.n_up + 1]] = dn[:, [s.n_up + 1, s.n_up]]
    sg3, la3 = o.log_psi(th, dn)
    assert np.allclose(la, la3, rtol=1e-5) and np.array_equal(sg, -sg3)
    op = r.copy(); op[:, [0, s.n_up]] = op[:, [s.n_up, 0]]           # opposite spins: no constraint
    assert not np.allclose(la, o.log_psi(th, op)[1], rtol=1e-5)


@pytest.mark.parametrize("kind,cusp", [(1, False), (1, True), (0, False)])
def test_param_gradient_finite_differences(kind, cusp):
    """K-4 (phase2-issue-body.md:243): analytic d log|psi| / d theta vs central differences < 1e-4."""
    s, net, th, r = _case(4, hidden=8, kind=kind, cusp=cusp, W=3)
    o = Oracle(s, net)
    O = o.grad_log_psi(th, r)
    rng = np.random.default_rng(3)
    for p in rng.choice(param_count(s, net), size=60, replace=False):
        d = np.zeros_like(th); d[p] = 1e-5
        fd = (o.log_psi(th + d, r)[1] - o.log_psi(th - d, r)[1]) / 2e-5
        assert np.max(np.abs(fd - O[:, p])) < 1e-4 * max(1.0, np.max(np.abs(fd)))


def test_cusp_slopes():
    """K-5 (phase3-issue-body.md:222-226): with only the cusp terms active, d log|psi| / d d_iA -> -Z_A + g_A(0)
    at a nucleus and the e-e cusp term has slope c_ij (0.5 same spin / 0.25 opposite, as the spec states)."""
    s = hydrogen_chain_system(4)
    net = OracleNet(hidden=(8, 8), use_cusp=True)
    th = init_theta(s, net, 0)
    u = unflatten(s, net, th)
    for k in ("a_ee", "a_en"):
        u[f"params/jastrow/{k}"][...] = 0.0                       # switch the Jastrow off
    o = Oracle(s, net)
    p = unflatten(s, net, th)
    base = np.array([[[0.9, 0.3, 0.2], [2.7, -0.4, 0.1], [4.1, 0.5, -0.3], [6.0, 0.2, 0.4]]])
    # electron 0 approaches nucleus 0 along z
    def C(rr):
        return o._cusps(p, rr, False)[0][0]
    h = 1e-6
    r1, r2 = base.copy(), base.copy()
    r1[0, 0] = [0, 0, h]; r2[0, 0] = [0, 0, 2 * h]
    slope = (C(r2) - C(r1)) / h
    W0, b0 = p["params/en_cusp/Dense_0/kernel"], p["params/en_cusp/Dense_0/bias"]
    W1, b1 = p["params/en_cusp/Dense_1/kernel"], p["params/en_cusp/Dense_1/bias"]
    g0 = (np.tanh(np.array([0.0, 0.0, 1.0]) @ W0 + b0) @ W1)[0] + b1[0]
    # other nuclei contribute a smooth term: compare against the analytic directional derivative instead
    _, g, _, _ = o._cusps(p, r1, True)
    assert abs(slope - g[0, 0, 2]) < 1e-4
    # isolate nucleus 0's own cusp: the singular part of the slope is -Z + g(0)
    far = g[0, 0, 2] - (-s.Z[0] + g0)
    assert abs(far) < 1.5                                          # remainder is the smooth far-nuclei part
    # e-e: electrons 0,1 (same spin) and 0,2 (opposite) approach coalescence
    for (i, j, c) in ((0, 1, net.cusp_same), (0, 2, net.cusp_diff)):
        ra, rb = base.copy(), base.copy()
        ra[0, j] = ra[0, i] + [0, 0, h]; rb[0, j] = rb[0, i] + [0, 0, 2 * h]
        d_en = lambda rr: o._cusps(p, rr, False)[0][0]
        # subtract the e-n part, which is smooth in r_ij here, by differencing the pair term directly
        b = abs(p["params/ee_cusp/b_ee"])
        pair = lambda d: c * d / (1 + b * d)
        assert abs((pair(2 * h) - pair(h)) / h - c) < 1e-5


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10."""
    def h(a):
        return [int(x) for x in a]
    assert h(philox.philox4x32_10(np.zeros(4, np.uint32), np.zeros(2, np.uint32))) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert h(philox.philox4x32_10(np.full(4, 0xffffffff, np.uint32), np.full(2, 0xffffffff, np.uint32))) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert h(philox.philox4x32_10(np.array([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], np.uint32),
                                  np.array([0xa4093822, 0x299f31d0], np.uint32))) == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    n, u, _ = philox.walker_streams(42, 7, np.arange(50000), 4)
    assert abs(n.mean()) < 0.01 and abs(n.std() - 1) < 0.01 and 0 <= u.min() and u.max() < 1
    n2, u2, _ = philox.walker_streams(42, 7, np.arange(100, 200), 4)
    assert np.array_equal(n2, n[100:200]) and np.array_equal(u2, u[100:200])     # keyed by global walker id


def test_estimator_identities():
    """vmc.py:104-135 restated: gradient from raw sums == centred two-pass form; backflow runs in the autodiff oracle."""
    s, net, th, r = _case(4, W=40)
    o = Oracle(s, net)
    g, mean, std, e, O = o.energy_and_gradient(th, r)
    n = len(e)
    g2 = 2 * ((e[:, None] * O).sum(0) / n - mean * O.sum(0) / n)
    assert np.allclose(g, g2, rtol=1e-10, atol=1e-12)
    assert np.isclose(std, np.std(e) / np.sqrt(n))
    sb, netb, thb, rb = _case(4, bf=True, W=4)
    a = AutodiffOracle(sb, netb)
    eb = a.local_energy(thb, rb)
    assert np.all(np.isfinite(eb))
    swp = rb.copy(); swp[:, [0, 1]] = swp[:, [1, 0]]
    s1, l1 = a.log_psi(thb, rb); s2, l2 = a.log_psi(thb, swp)
    assert np.allclose(l1, l2, rtol=1e-8) and np.array_equal(s1, -s2)     # backflow keeps antisymmetry


@pytest.mark.parametrize("n_atoms,cusp", [(2, False), (4, False), (6, True)])
def test_backflow_analytic_equals_autodiff_fp64(n_atoms, cusp):
    """K-8 for the backflow ansatz: orbital-Hessian chain rule (Appendix C.3) == nested autodiff."""
    from oracle.nqmc_backflow import BackflowOracle
    s, net, th, r = _case(n_atoms, cusp=cusp, bf=True, W=6)
    o, a = BackflowOracle(s, net), AutodiffOracle(s, net)
    sg, la = o.log_psi(th, r)
    sg2, la2 = a.log_psi(th, r)
    assert np.array_equal(sg, sg2) and np.max(np.abs(la - la2)) < 1e-10
    e, _, _, _ = o.local_energy(th, r)
    e2 = a.local_energy(th, r)
    assert np.max(np.abs(e - e2) / np.maximum(1, np.abs(e2))) < 1e-9
    O, O2 = o.grad_log_psi(th, r), a.grad_log_psi(th, r)
    assert np.max(np.abs(O - O2) / np.maximum(1, np.abs(O2))) < 1e-9


def test_sr_oracle_matches_the_executed_spec_block():
    """oracle/sr.py against tests/golden/ref_sr_golden.npz (made by executing .github/phase6-issue-body.md:140-181)."""
    import os
    from oracle.sr import sr_update
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sr_golden.npz"))
    for c in "abc":
        lr, reg = g[f"{c}_lr_reg"]
        d = sr_update(g[f"{c}_grads"], g[f"{c}_energies"], lr, reg)
        assert np.allclose(d, g[f"{c}_update"], rtol=1e-9, atol=1e-12)
