"""-m gpu: the BASELINE configs beyond configs[1], at their per-GPU full sizes, through size-independent
properties (the oracle is too slow at these sizes) plus an oracle check on a sampled subset.

  configs[2]  H6 chain (3up/3dn), backflow + cusp, H=128: 262,144 walkers over 8 GPUs = 32,768 per GPU
  configs[3]  H4 dissociation sweep, 16 geometries R = 1.0..4.0 as independent walker sets (16,384 each)
  configs[4]  H10 chain (5up/5dn), H=128: 1,048,576 walkers over 8 GPUs = 131,072 per GPU
"""
import numpy as np
import pytest

from oracle.layout import OracleNet, hydrogen_chain_system, init_theta

from _util import T, make_context, make_oracle, rel_err

pytestmark = pytest.mark.gpu


def _setup(n_atoms, hidden, W, cusp=False, bf=False, seed=0):
    import torch
    s = hydrogen_chain_system(n_atoms)
    net = OracleNet(hidden=(hidden, hidden), use_cusp=cusp, use_backflow=bf)
    theta = init_theta(s, net, seed + 1).astype(np.float32)
    ctx = make_context(s, net)
    ctx.set_params(theta)
    rng = np.random.default_rng(seed)
    r = (s.R[np.arange(s.n_elec) % s.n_atoms][None] + 0.5 * rng.standard_normal((W, s.n_elec, 3))).astype(np.float32)
    soa = T(ctx.to_soa(r))
    _, la = ctx.log_psi(soa)
    logp = (2 * T(la)).contiguous()
    for t in range(20):                                        # equilibrate on the device
        ctx.mh_step(soa, logp, 0.3, seed=7, step=t)
    return s, net, theta, ctx, soa, logp


@pytest.mark.parametrize("name,n_atoms,hidden,W,cusp,bf", [
    ("H4-65536", 4, 64, 65536, False, False),               # the benchmarked configuration on the benchmarked (tcgen05) path
    ("H6-bf-cusp-32768", 6, 128, 32768, True, True),
    ("H10-131072", 10, 128, 131072, False, False),
])
def test_full_size_properties(name, n_atoms, hidden, W, cusp, bf):
    import torch
    s, net, theta, ctx, soa, logp = _setup(n_atoms, hidden, W, cusp, bf)
    dev = ctx.dev()
    e_l = torch.empty(W, dtype=torch.float32, device=dev)
    hdr = torch.empty(8, dtype=torch.float64, device=dev)
    gsum = torch.empty(ctx.P, dtype=torch.float64, device=dev)
    r_before = soa.clone()
    ctx.walker_step(soa, logp, 0.3, e_l, hdr, gsum, seed=11, step=5)
    h = hdr.cpu().numpy()
    e = e_l.cpu().numpy()
    fin = np.isfinite(e)
    # (1) header is the checksum of the per-walker outputs
    assert h[0] + h[4] == W and h[0] == fin.sum()
    assert abs(h[1] - e[fin].astype(np.float64).sum()) <= 1e-9 * np.abs(e[fin]).sum()
    moved = (soa != r_before).any(dim=0).sum().item()
    assert moved == h[3] and 0.2 * W < moved < 0.98 * W          # accepted walkers moved, the others did not
    # (2) antisymmetry on the full batch: swap two same-spin electrons => log|psi|, E_L equal, sign flips
    sg, la = ctx.log_psi(soa)
    sw = soa.clone()
    sw[0:3], sw[3:6] = soa[3:6].clone(), soa[0:3].clone()
    sg2, la2 = ctx.log_psi(sw)
    la_c, la2_c = la.cpu().numpy(), la2.cpu().numpy()
    assert np.quantile(rel_err(la2_c, la_c), 0.99) <= 2e-5
    assert (sg.cpu().numpy() == -sg2.cpu().numpy()).mean() >= 0.999
    e_sw = ctx.local_energy(sw).cpu().numpy()
    e_now = ctx.local_energy(soa).cpu().numpy()
    assert np.median(rel_err(e_sw, e_now)) <= 1e-5 and np.quantile(rel_err(e_sw, e_now), 0.95) <= 1e-3
    assert np.array_equal(e_now, e)                               # E_L is a pure function of (theta, r): idempotent
    # (3) sharding invariance: a sub-block evaluated on its own gives bit-identical per-walker results
    lo, n = 4096, 8192
    sub = soa[:, lo:lo + n].contiguous()
    assert np.array_equal(ctx.local_energy(sub).cpu().numpy(), e_now[lo:lo + n])
    # (4) linearity of the weighted gradient accumulation
    rng = np.random.default_rng(1)
    w1 = torch.as_tensor(rng.standard_normal(W).astype(np.float32)).to(dev)
    w2 = torch.as_tensor(rng.standard_normal(W).astype(np.float32)).to(dev)
    g1, g2, g12 = ctx.grad_accum(soa, w1), ctx.grad_accum(soa, w2), ctx.grad_accum(soa, (w1 + 2 * w2).contiguous())
    lin = g1.cpu().numpy() + 2 * g2.cpu().numpy()
    got = g12.cpu().numpy()
    # (random-sign weights over 1e5 walkers cancel heavily: fp32 partial sums, so 1e-4 rather than 1e-5)
    assert np.median(np.abs(got - lin) / np.maximum(np.abs(lin), 1e-3)) <= 1e-4
    # (5) centred estimator: gsum == sum_w (E_w - mean) O_w == grad_accum(weights = E - mean)
    mean = h[1] / h[0]
    wc = torch.where(torch.isfinite(e_l), e_l - float(np.float32(mean)), torch.zeros_like(e_l)).contiguous()
    gc = ctx.grad_accum(soa, wc).cpu().numpy()
    gs = gsum.cpu().numpy()
    assert np.median(np.abs(gc - gs) / np.maximum(np.abs(gs), 1e-3)) <= 1e-4
    # (6) oracle on a sampled subset
    idx = rng.choice(W, size=48, replace=False)
    r_aos = ctx.to_aos(soa).cpu().numpy()[idx].astype(np.float64)
    o = make_oracle(s, net)
    e_ref, _, sg_ref, la_ref = o.local_energy(theta.astype(np.float64), r_aos)
    assert np.median(rel_err(e[idx], e_ref)) <= 1e-4
    assert np.median(rel_err(la_c[idx], la_ref)) <= 1e-5
    # (7) reverse pass at full launch size against the oracle: weights that are non-zero on the subset only
    if not bf:
        wsub = np.zeros(W, dtype=np.float32)
        wsub[idx] = rng.standard_normal(idx.size).astype(np.float32)
        gsub = ctx.grad_accum(soa, torch.as_tensor(wsub).to(dev)).cpu().numpy()
        O = o.grad_log_psi(theta.astype(np.float64), r_aos)
        ref = (wsub[idx].astype(np.float64)[:, None] * O).sum(0)
        scale = (np.abs(wsub[idx]).astype(np.float64)[:, None] * np.abs(O)).sum(0)
        err = np.abs(gsub - ref) / np.maximum(scale, 1e-3)
        assert np.median(err) <= 1e-5 and np.quantile(err, 0.99) <= 1e-3, (np.median(err), np.quantile(err, 0.99))


def test_h4_dissociation_sweep_16_geometries():
    import torch
    from nqmc_b200.sweep import GeometrySweep
    bonds = np.linspace(1.0, 4.0, 16)
    sw = GeometrySweep(bond_lengths=bonds, walkers_per_geometry=16384, hidden=(64, 64), step_size=0.5, seed=3)
    assert len(sw.states) == 16
    sw.burn_in(10)
    sw.synchronize()
    before = [(st.r.clone(st.ctx.stream), st.logp.clone(st.ctx.stream)) for st in sw.states]   # each geometry runs on its own stream
    sw.step()
    sw.synchronize()
    res = sw.results()
    for g, (st, out) in enumerate(zip(sw.states, res)):
        assert out["bond_length"] == pytest.approx(bonds[g])
        assert out["v_nn"] == pytest.approx(13.0 / (3.0 * bonds[g]), rel=1e-12)
        assert np.isfinite(out["energy"]) and out["std"] > 0 and np.all(np.isfinite(out["gradient"]))
        if g in (0, 7, 15):
            # independence: re-running geometry g alone from the same state reproduces its numbers -- E_L, walkers and
            # the statistics header bit for bit, the gradient sums to fp32-atomic ordering
            ctx = st.ctx
            r2, lp2 = before[g]
            e2, h2, g2 = ctx.empty(sw.W), ctx.empty(8, np.float64), ctx.empty(ctx.P, np.float64)
            ctx.walker_step(r2, lp2, sw.step_size, e2, h2, g2, seed=sw.seed + 1000 * g, step=(1 << 20) + sw.t)
            ctx.synchronize()
            assert np.array_equal(e2.cpu().numpy(), st.e_l.cpu().numpy())
            assert np.array_equal(r2.cpu().numpy(), st.r.cpu().numpy())
            assert np.array_equal(h2.cpu().numpy(), st.hdr.cpu().numpy())
            ga, gb = g2.cpu().numpy(), st.gsum.cpu().numpy()
            assert np.max(np.abs(ga - gb)) <= 1e-6 * np.max(np.abs(gb))
    # geometry 0 and 15 against the oracle on a subset (different nuclei / V_nn per context)
    for g in (0, 15):
        st = sw.states[g]
        s = hydrogen_chain_system(4, float(bonds[g]))
        net = OracleNet(hidden=(64, 64))
        from nqmc_b200.params import ravel
        th = ravel(st.params).astype(np.float64)
        r_aos = st.ctx.to_aos(st.r).cpu().numpy()[:64].astype(np.float64)
        e_ref, parts, _, _ = make_oracle(s, net).local_energy(th, r_aos)
        assert parts["V_nn"] == pytest.approx(13.0 / (3.0 * bonds[g]))
        assert np.median(rel_err(st.e_l.cpu().numpy()[:64], e_ref)) <= 1e-4
    # energies differ between geometries (they are not accidentally sharing a context)
    en = np.array([o["energy"] for o in res])
    assert len(np.unique(np.round(en, 6))) == 16


@pytest.mark.parametrize("case", [dict(n_atoms=4, hidden=128), dict(n_atoms=6, hidden=128, use_cusp=True, use_backflow=True)],
                         ids=["H4-h128", "H6-h128-bf-cusp"])
def test_reverse_pass_from_parked_activations_equals_value_sweep(case):
    """Width 128: the fused step's reverse pass reads h1 / h2 / features parked by the 5-channel energy pass
    (bwd128_seed_kernel); an energy call between the two phases invalidates them and the reverse pass runs its own value
    sweep.  Both must give the same gradient sums (tensor-core kernels on both sides: 3xTF32 vs FP32 layer 1)."""
    from _util import make_case, make_context
    W = 1000
    s, net, theta, r = make_case(W=W, burn=5, **case)
    ctx = make_context(s, net)
    ctx.set_params(theta)
    out = []
    for invalidate in (False, True):
        soa = ctx.to_soa(r)
        _, la = ctx.log_psi(soa)
        logp = ctx.upload((2 * la.numpy()).astype(np.float32))
        e, hdr, g = ctx.empty(W), ctx.empty(8, np.float64), ctx.empty(ctx.P, np.float64)
        ctx.walker_step_a(soa, logp, 0.3, e, hdr, seed=11, step=0)
        if invalidate:
            ctx.local_energy(soa)                                      # any energy pass drops the parked activations
        ctx.walker_step_b(soa, e, hdr, g)
        out.append((e.numpy(ctx.stream).copy(), g.numpy(ctx.stream).copy()))
    assert np.array_equal(out[0][0], out[1][0])
    ga, gb = out
    scale = max(1.0, float(np.max(np.abs(gb[1]))))
    assert np.max(np.abs(ga[1] - gb[1])) <= 2e-5 * scale, (np.max(np.abs(ga[1] - gb[1])), scale)
    assert np.any(ga[1] != gb[1])                                       # the two routes really are different kernels
