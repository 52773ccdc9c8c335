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
