"""CPU: the oracle's two restatements against fixtures produced by executing the REFERENCE'S OWN SOURCE
(hamiltonian/molecular.py, wavefunction/simple.py, sampler/metropolis.py, optimizer/vmc.py, and the [SPEC]
code blocks of .github/phase{2,3}-issue-body.md) under oracle/_ref_shim -- see
tests/golden/make_ref_hotpath_golden.py.  This is what pins the oracle; the -m gpu twin
(test_gpu_ref_golden.py) checks the CUDA path against the same files.

Fixture values are float64 (jax_enable_x64); both restatements must reproduce them to ~1e-9."""
import os

import numpy as np
import pytest

from oracle.layout import leaf_table, param_count
from oracle.nqmc_autodiff import AutodiffOracle

from _golden import CASE_NAMES, MANIFEST, load
from _util import make_oracle


def close(a, b, tol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) if a.size else 0.0
    assert err <= tol, f"{what}: max rel err {err:.3e} > {tol}"


@pytest.fixture(scope="module", params=CASE_NAMES)
def gold(request):
    meta, a, s, net = load(request.param)
    return dict(meta=meta, a=a, s=s, net=net, o=make_oracle(s, net), th=a["theta"].astype(np.float64))


def test_parameter_tree_matches_flax_naming(gold):
    """Leaf paths / shapes / order produced by the flax stand-in running the reference's modules
    == oracle/layout.py's table (and hence the CUDA library's, test_gpu_parity.py::test_layout_matches_oracle)."""
    ref = [(p, tuple(sh)) for p, sh in gold["meta"]["leaves"]]
    ours = [(p, tuple(sh)) for p, sh, _ in leaf_table(gold["s"], gold["net"])]
    assert ours == ref
    assert param_count(gold["s"], gold["net"]) == gold["meta"]["P"] == gold["a"]["theta"].size


def test_log_psi_and_sign(gold):
    a = gold["a"]
    sg, la = gold["o"].log_psi(gold["th"], a["r"])
    close(la, a["log_psi"], 1e-10, "log|psi|")
    if "sign" in a:
        assert np.array_equal(sg, a["sign"])


def test_local_energy_parts(gold):
    a = gold["a"]
    e, parts, _, _ = gold["o"].local_energy(gold["th"], a["r"])
    close(parts["T"], a["T"], 2e-8, "T")
    close(parts["V_en"], a["V_en"], 1e-12, "V_en")
    close(parts["V_ee"], a["V_ee"], 1e-12, "V_ee")
    close(e, a["E_L"], 2e-8, "E_L")
    assert abs(parts["V_nn"] - gold["meta"]["V_nn"]) < 1e-12


def test_mh_step(gold):
    a, m = gold["a"], gold["meta"]
    r2, lp2, acc, ratio = gold["o"].mh_step(gold["th"], a["r"], 2.0 * a["log_psi"], a["mh_normals"], a["mh_uniforms"],
                                            m["step_size"])
    close(ratio, a["mh_ratio"], 1e-9, "log accept ratio")
    assert np.array_equal(acc.astype(np.int8), a["mh_accept"])
    close(r2, a["mh_positions_out"], 1e-12, "positions")
    close(lp2, a["mh_logp_out"], 1e-9, "log_prob")


def test_sample_burn_thin_and_acceptance(gold):
    """MetropolisSampler.sample semantics (metropolis.py:174-218): n_burn steps, counters reset, n_samples*n_thin
    steps, every n_thin-th kept starting with the FIRST (all_positions[::n_thin]), rate = mean(n_accepted)/n_steps."""
    a, m, o = gold["a"], gold["meta"], gold["o"]
    r, lp = a["r"], 2.0 * a["log_psi"]
    nb, ns, nt = m["n_burn"], m["n_samples"], m["n_thin"]
    kept, nacc = [], np.zeros(r.shape[0])
    for t in range(nb + ns * nt):
        r, lp, acc, _ = o.mh_step(gold["th"], r, lp, a["smp_normals"][t], a["smp_uniforms"][t], m["step_size"])
        if t >= nb:
            nacc += acc
            if (t - nb) % nt == 0:
                kept.append(r.copy())
    close(np.stack(kept), a["smp_samples"], 1e-12, "samples")
    close(nacc, a["smp_final_n_accepted"], 0, "n_accepted")
    close(nacc.mean() / (ns * nt), a["smp_acc_rate"], 1e-12, "acceptance rate")
    close(lp, a["smp_final_logp"], 1e-9, "final log_prob")


def test_gradient_estimator(gold):
    """compute_gradient / estimate_energy (vmc.py:49-135) on the sampled configurations."""
    a, o, th = gold["a"], gold["o"], gold["th"]
    flat = a["smp_samples"].reshape(-1, *a["smp_samples"].shape[2:])
    e, _, _, _ = o.local_energy(th, flat)
    O = o.grad_log_psi(th, flat)
    g = 2.0 * np.mean((e - e.mean())[:, None] * O, axis=0)
    scale = max(1.0, np.abs(a["grad"]).max())
    assert np.max(np.abs(g - a["grad"])) <= 1e-8 * scale
    close(e.mean(), a["grad_mean_energy"], 1e-9, "mean energy")
    close(e.std() / np.sqrt(e.size), a["grad_energy_std"], 1e-8, "energy std")
    close(e.mean(), a["est_mean"], 1e-9, "estimate_energy mean")
    close(e.std() / np.sqrt(e.size), a["est_std"], 1e-8, "estimate_energy std")
    if "mean_O" in a:
        assert np.max(np.abs(O.mean(0) - a["mean_O"])) <= 1e-9 * max(1.0, np.abs(a["mean_O"]).max())


def test_autodiff_restatement(gold):
    """Restatement (i) (torch.func nested autodiff, the CPU baseline bench.py times) against the same fixtures."""
    a = gold["a"]
    ad = AutodiffOracle(gold["s"], gold["net"])
    n = min(8, a["r"].shape[0])
    sg, la = ad.log_psi(gold["th"], a["r"][:n])
    close(la, a["log_psi"][:n], 1e-10, "log|psi|")
    close(ad.local_energy(gold["th"], a["r"][:n]), a["E_L"][:n], 2e-8, "E_L")


def test_float32_reference_is_within_the_stated_tolerances(gold):
    """The reference's shipped precision (float32) vs its own float64 evaluation: the yardstick the CUDA path is
    held to.  Documents that the north_star tolerances are about arithmetic, not the algorithm."""
    a = gold["a"]
    err_l = np.abs(a["f32_log_psi"] - a["log_psi"]) / np.maximum(1, np.abs(a["log_psi"]))
    err_e = np.abs(a["f32_E_L"] - a["E_L"]) / np.maximum(1, np.abs(a["E_L"]))
    assert np.median(err_l) <= 1e-5 and np.median(err_e) <= 1e-4


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/nqmc"), reason="reference tree not present (GPU box)")
def test_fixture_regenerates_from_the_reference_source(tmp_path):
    """In the build container: re-run the generator for one case and require the committed fixture bit for bit,
    and the recorded sha256 of every reference file to match what is on disk."""
    import hashlib
    import subprocess
    import sys
    for rel, h in MANIFEST["reference_files"].items():
        assert hashlib.sha256(open(os.path.join("/root/reference", rel), "rb").read()).hexdigest() == h, rel
    here = os.path.dirname(os.path.abspath(__file__))
    code = ("import sys, numpy as np; sys.argv=['x']; sys.path.insert(0, %r); import make_ref_hotpath_golden as g; "
            "mk_mol, mk_wf, kw, d = g.CASES['simple_h2_silu3']; mol = mk_mol(); "
            "out, meta = g.run_case('simple_h2_silu3', mol, lambda: mk_wf(mol), **kw); np.savez(%r, **out)"
            % (os.path.join(here, "golden"), str(tmp_path / "o.npz")))
    subprocess.check_call([sys.executable, "-c", code])
    new = np.load(tmp_path / "o.npz")
    old = np.load(os.path.join(here, "golden", "ref_hotpath_simple_h2_silu3.npz"))
    assert sorted(new.files) == sorted(old.files)
    for k in old.files:
        assert np.array_equal(new[k], old[k]), k
