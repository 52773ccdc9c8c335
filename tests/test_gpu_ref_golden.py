"""-m gpu: the CUDA path (through the C ABI) against fixtures produced by EXECUTING THE REFERENCE'S OWN SOURCE
(tests/golden/make_ref_hotpath_golden.py: hamiltonian/molecular.py, wavefunction/simple.py, sampler/metropolis.py,
optimizer/vmc.py and the [SPEC] code blocks, run under oracle/_ref_shim with jax_enable_x64).

Tolerances (BASELINE.json north_star): log|psi| <= 1e-5 relative, E_L <= 1e-4 relative, with relative error
|d| / max(1, |ref|); sign equal; accept/reject identical outside the band |log u - dlogp| <= 1e-4 max(1, |dlogp|).
The fixtures also hold the reference evaluated in its shipped float32 (``f32_*``): where a walker sits close to a
node and float32 itself cannot reach the tolerance, the CUDA error must stay within 4x of that float32 error.
"""
import numpy as np
import pytest

from _golden import CASE_NAMES, load
from _util import make_context, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=CASE_NAMES)
def gold(request):
    meta, a, s, net = load(request.param)
    ctx = make_context(s, net)
    ctx.set_params(a["theta"])
    return dict(meta=meta, a=a, s=s, net=net, ctx=ctx)


def _within(err, tol, err32, what):
    """every walker within tol, or within 4x of what the reference's own float32 evaluation achieves there"""
    bad = err > np.maximum(tol, 4.0 * err32)
    assert not bad.any(), f"{what}: {int(bad.sum())} walkers off; worst {err.max():.2e} (float32 reference {err32.max():.2e})"
    assert np.median(err) <= tol, f"{what}: median {np.median(err):.2e}"


def test_layout_is_the_flax_tree(gold):
    ours = [(p, tuple(x for x in sh if x != 1) or ()) for p, sh, _ in gold["ctx"].leaves()]
    ref = [(p, tuple(x for x in sh if x != 1) or ()) for p, sh in gold["meta"]["leaves"]]
    assert [p for p, _ in ours] == [p for p, _ in ref]
    assert [int(np.prod(s)) for _, s in ours] == [int(np.prod(s)) for _, s in ref]
    assert gold["ctx"].P == gold["meta"]["P"]


def test_log_psi_and_sign(gold):
    a, ctx = gold["a"], gold["ctx"]
    sg, la = ctx.log_psi(ctx.to_soa(a["r"].astype(np.float32)))
    sg, la = sg.cpu().numpy(), la.cpu().numpy()
    _within(rel_err(la, a["log_psi"]), 1e-5, rel_err(a["f32_log_psi"], a["log_psi"]), "log|psi|")
    if "sign" in a:
        assert np.array_equal(sg, a["sign"].astype(np.float32))
    else:
        assert np.all(sg == 1.0)


def test_local_energy(gold):
    a, ctx = gold["a"], gold["ctx"]
    e, parts = ctx.local_energy(ctx.to_soa(a["r"].astype(np.float32)), want_parts=True)
    e, parts = e.cpu().numpy(), parts.cpu().numpy()
    _within(rel_err(e, a["E_L"]), 1e-4, rel_err(a["f32_E_L"], a["E_L"]), "E_L")
    _within(rel_err(parts[0], a["T"]), 1e-4, rel_err(a["f32_T"], a["T"]), "T")
    assert np.max(rel_err(parts[1], a["V_en"])) <= 1e-5
    assert np.max(rel_err(parts[2], a["V_ee"])) <= 1e-5


def test_mh_step(gold):
    import torch
    a, m, ctx = gold["a"], gold["meta"], gold["ctx"]
    dev = ctx.dev()
    r = ctx.to_soa(a["r"].astype(np.float32))
    logp = torch.as_tensor((2.0 * a["log_psi"]).astype(np.float32)).to(dev)
    W = a["r"].shape[0]
    acc = torch.zeros(W, dtype=torch.uint8, device=dev)
    nacc = torch.zeros(W, dtype=torch.float32, device=dev)
    ctx.mh_step(r, logp, m["step_size"], normals=ctx.to_soa(a["mh_normals"].astype(np.float32)),
                uniforms=torch.as_tensor(a["mh_uniforms"].astype(np.float32)).to(dev), accept=acc, n_accepted=nacc)
    acc = acc.cpu().numpy()
    with np.errstate(divide="ignore"):
        logu = np.log(a["mh_uniforms"])
    band = np.abs(logu - a["mh_ratio"]) <= 1e-4 * np.maximum(1.0, np.abs(a["mh_ratio"]))
    assert np.array_equal(acc[~band], a["mh_accept"][~band]), "accept/reject differs away from the threshold"
    assert band.mean() < 0.2
    same = acc == a["mh_accept"]
    r_new = ctx.to_aos(r).cpu().numpy()
    assert np.max(np.abs(r_new[same] - a["mh_positions_out"][same])) <= 1e-6
    assert np.max(rel_err(logp.cpu().numpy()[same], a["mh_logp_out"][same])) <= 2e-5
    assert np.array_equal(nacc.cpu().numpy(), acc.astype(np.float32))


def test_gradient_estimator(gold):
    """2 <(E - <E>) O> on the reference's sampled configurations (optimizer/vmc.py:104-135), via E_L -> header ->
    centred reverse pass, exactly what VMCOptimizer.step runs."""
    import torch
    a, ctx = gold["a"], gold["ctx"]
    flat = a["smp_samples"].reshape(-1, *a["smp_samples"].shape[2:]).astype(np.float32)
    r = ctx.to_soa(flat)
    e = ctx.local_energy(r)
    hdr = ctx.energy_stats(e)
    gsum = torch.empty(ctx.P, dtype=torch.float64, device=ctx.dev())
    ctx.walker_step_b(r, e, hdr, gsum)
    h = hdr.cpu().numpy()
    g = 2.0 * gsum.cpu().numpy() / h[0]
    mean = h[1] / h[0]
    assert abs(mean - a["grad_mean_energy"]) <= 1e-4 * max(1.0, abs(a["grad_mean_energy"]))
    std = np.sqrt(max(h[2] / h[0] - mean * mean, 0.0)) / np.sqrt(h[0])
    assert abs(std - a["grad_energy_std"]) <= 1e-3 * max(1.0, abs(a["grad_energy_std"]))
    scale = np.abs(a["grad"]).max()
    err = np.abs(g - a["grad"]) / scale
    if "f32_grad" in a:      # the reference in its own float32
        err32 = np.abs(a["f32_grad"] - a["grad"]) / scale
        assert err.max() <= max(1e-4, 4.0 * err32.max()), (err.max(), err32.max())
    else:
        assert err.max() <= 1e-3, err.max()
    assert np.median(err) <= 1e-5


# ------------------------------------------------------------------------------------------------------------------
# the reference-facing Python API (nqmc_b200.{wavefunction,sampler,optimizer}) against the same fixtures
def _host_objects(meta, a):
    """(molecule, wavefunction, params) of the drop-in package for a fixture."""
    import nqmc_b200 as nqmc
    from nqmc_b200.params import unravel
    from nqmc_b200.systems.molecule import Molecule
    from nqmc_b200.wavefunction import AntisymmetricWavefunction, SimpleWavefunction
    g = meta["molecule"]
    n_el = sum(g["spin"])
    mol = Molecule.from_atoms(["H"] * len(g["charges"]), np.asarray(g["positions"]), spin_polarization=g["spin"][0] - g["spin"][1])
    assert mol.spin == tuple(g["spin"]) and mol.n_electrons == n_el
    if meta["kind"] == "simple":
        wf = SimpleWavefunction(hidden_dims=tuple(meta["hidden"]), activation=meta["activation"],
                                envelope_decay=meta.get("envelope_decay", 1.0), nuclear_positions=mol.positions)
        like = wf.init(nqmc.random.PRNGKey(0), np.zeros((n_el, 3)))
    else:
        wf = AntisymmetricWavefunction(g["spin"][0], g["spin"][1], tuple(meta["hidden"]), use_cusp=meta["use_cusp"],
                                       use_backflow=meta["use_backflow"], molecule=mol)
        like = wf.init(nqmc.random.PRNGKey(0))
    return mol, wf, unravel(a["theta"], like)


def test_host_api_sample_burn_thin(gold):
    """MetropolisSampler.sample (metropolis.py:149-218) through the drop-in class, on the reference's streams."""
    import nqmc_b200 as nqmc
    from nqmc_b200.sampler import MetropolisSampler, SamplerState
    a, m = gold["a"], gold["meta"]
    mol, wf, params = _host_objects(m, a)
    wf.context(mol)
    W = a["r"].shape[0]
    st0 = SamplerState(a["r"].astype(np.float32), (2.0 * a["log_psi"]).astype(np.float32), np.zeros(W, np.float32), 0)
    smp = MetropolisSampler(step_size=m["step_size"])
    samples, st, rate = smp.sample(nqmc.random.PRNGKey(1), wf.log_prob_fn(params), st0, n_samples=m["n_samples"],
                                   n_burn=m["n_burn"], n_thin=m["n_thin"], normals=a["smp_normals"], uniforms=a["smp_uniforms"])
    assert samples.shape == a["smp_samples"].shape and st.n_steps == m["n_samples"] * m["n_thin"]
    # walkers whose every accept decision matched end at the same place; a decision may flip only inside the band
    agree = np.all(np.abs(samples - a["smp_samples"]) <= 2e-6, axis=(0, 2, 3))
    assert agree.mean() >= 0.9, agree.mean()
    assert np.array_equal(np.asarray(st.n_accepted)[agree], a["smp_final_n_accepted"][agree])
    assert abs(rate - float(a["smp_acc_rate"])) <= 2.0 / W + (1 - agree.mean())
    assert np.max(rel_err(np.asarray(st.log_prob)[agree], a["smp_final_logp"][agree])) <= 5e-5


VMC_CASES = [n for n in CASE_NAMES if "vmc" in load(n)[0]]


@pytest.mark.parametrize("name", VMC_CASES)
@pytest.mark.parametrize("device_update", [False, True])
def test_host_api_vmc_step_trajectory(name, device_update):
    """VMCOptimizer.step (vmc.py:217-282) -- sample, gradient, optax clip + Adam, metrics -- against the parameter
    trajectory the reference's own class produced on the same streams."""
    import nqmc_b200 as nqmc
    from nqmc_b200.optimizer import VMCOptimizer
    from nqmc_b200.optimizer.vmc import VMCState, _opt_init
    from nqmc_b200.params import ravel
    from nqmc_b200.sampler import SamplerState
    meta, a, s, net = load(name)
    mol, wf, params = _host_objects(meta, a)
    v = meta["vmc"]
    W = a["vmc_init_positions"].shape[0]
    opt = VMCOptimizer(learning_rate=v["learning_rate"], n_samples=v["n_samples"], n_chains=W, n_burn=v["n_burn"],
                       step_size=meta["step_size"], clip_grad=v["clip_grad"], device_update=device_update)
    st = VMCState(params, _opt_init(params), SamplerState(a["vmc_init_positions"].astype(np.float32),
                                                         a["vmc_init_logp"].astype(np.float32), np.zeros(W, np.float32), 0), 0)
    key = nqmc.random.PRNGKey(0)
    for i in range(v["steps"]):
        st, m = opt.step(key, st, wf, mol, normals=a[f"vmc_normals_{i}"], uniforms=a[f"vmc_uniforms_{i}"])
        ref = dict(zip(v["metric_keys"], a["vmc_metrics"][i]))
        th = ravel(st.params).astype(np.float64)
        d = np.abs(th - a["vmc_theta"][i])
        tight = i == 0       # later steps: an accept decision may flip once theta differs in the 7th digit
        assert abs(m["energy"] - ref["energy"]) <= (1e-4 if tight else 5e-2) * max(1.0, abs(ref["energy"])), (i, m, ref)
        if tight:
            assert abs(m["energy_std"] - ref["energy_std"]) <= 1e-3 * max(1.0, ref["energy_std"])
            assert abs(m["acceptance_rate"] - ref["acceptance_rate"]) <= 2.0 / W
            assert abs(m["grad_norm"] - ref["grad_norm"]) <= 1e-3 * max(1.0, ref["grad_norm"])
            # Adam's first step is -lr * sign(g): a parameter whose gradient is zero to rounding may land lr away
            assert np.mean(d <= 1e-5) >= 0.99, np.mean(d <= 1e-5)
        assert np.median(d) <= (1e-6 if tight else 1e-3)
        assert d.max() <= 2.5 * v["learning_rate"] * (i + 1)
    assert st.step == v["steps"] and st.opt_state["count"] == v["steps"]
