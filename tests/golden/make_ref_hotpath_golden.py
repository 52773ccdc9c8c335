"""Generates tests/golden/ref_hotpath_*.npz by EXECUTING THE REFERENCE'S OWN SOURCE FILES.

Run in the build container only (``/root/reference`` does not travel to the GPU box):

    python tests/golden/make_ref_hotpath_golden.py            # all cases
    python tests/golden/make_ref_hotpath_golden.py simple_h2  # one case

How: ``oracle/_ref_shim`` provides import-name stand-ins for jax / flax.linen / optax over
torch.func (none of the three is installable here).  With it and ``/root/reference/src`` on
``sys.path``, ``import nqmc`` loads the reference package unmodified, and this script calls

    nqmc.wavefunction.simple.SimpleWavefunction.{init,__call__}      simple.py:28-176
    nqmc.hamiltonian.molecular.{kinetic_energy, electron_nuclear_potential,
        electron_electron_potential, local_energy, local_energy_batched}  molecular.py:36-197
    nqmc.sampler.metropolis.MetropolisSampler.{init_state, step, sample}   metropolis.py:56-218
    nqmc.optimizer.vmc.{compute_gradient, estimate_energy, VMCOptimizer.{init,step}}  vmc.py:49-282

For the Slater / Jastrow / cusp / backflow ansatz the reference has no importable module; its
definition is the [SPEC] starter code in ``.github/phase2-issue-body.md`` / ``phase3-issue-body.md``.
Those python blocks are EXTRACTED FROM THE MARKDOWN BY LINE RANGE and exec'd (slater 38-65,
orbitals 78-113, Jastrow 125-152; e-n cusp 50-83, e-e cusp 95-122, backflow 134-169).  Only the
top-level composition, which the spec writes in a form that does not run (phase3:180-219 mixes
``setup`` with ``@nn.compact`` and calls NumPy-only Molecule methods on traced arrays; phase4:97-111
``compute_slater``), is restated here in ``SpecAntisymmetric`` -- 25 lines, same order, same names.
The result is wrapped as a reference ``Wavefunction`` and pushed through the reference's real
``local_energy`` / ``MetropolisSampler`` / ``compute_gradient``.

Primary values are produced with ``jax_enable_x64`` (the algorithm without rounding noise); a
second pass in float32 (the reference's shipped precision) is stored under ``f32_*`` keys as the
yardstick for "fp32 arithmetic".  Random draws are recorded from the stand-in's tape.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "_ref_shim"), os.path.join(REF, "src")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402
from flax import linen as nn  # noqa: E402
from typing import Tuple  # noqa: E402

import nqmc  # noqa: E402  (the reference package itself)
from nqmc.hamiltonian.molecular import (electron_electron_potential, electron_nuclear_potential,  # noqa: E402
                                        kinetic_energy, local_energy_batched)
from nqmc.optimizer.vmc import VMCOptimizer, compute_gradient, estimate_energy  # noqa: E402
from nqmc.sampler.metropolis import MetropolisSampler  # noqa: E402
from nqmc.systems.molecule import Molecule, hydrogen_atom, hydrogen_chain, hydrogen_molecule  # noqa: E402
from nqmc.wavefunction.base import Wavefunction  # noqa: E402
from nqmc.wavefunction.simple import SimpleWavefunction  # noqa: E402

assert nqmc.__file__.startswith(REF), nqmc.__file__

SPEC_BLOCKS = [(".github/phase2-issue-body.md", 38, 65), (".github/phase2-issue-body.md", 78, 113),
               (".github/phase2-issue-body.md", 125, 152), (".github/phase3-issue-body.md", 50, 83),
               (".github/phase3-issue-body.md", 95, 122), (".github/phase3-issue-body.md", 134, 169)]


def load_spec_namespace():
    ns = {"jax": jax, "jnp": jnp, "nn": nn, "Tuple": Tuple}
    for rel, lo, hi in SPEC_BLOCKS:
        lines = open(os.path.join(REF, rel), encoding="utf-8").read().split("\n")
        assert lines[lo - 2].startswith("```python") and lines[hi].startswith("```"), (rel, lo, hi)
        exec(compile("\n".join(lines[lo - 1:hi]), f"{rel}:{lo}-{hi}", "exec"), ns)
    return ns


SPEC = load_spec_namespace()


class SpecAntisymmetric(nn.Module):
    """Composition of the extracted [SPEC] modules; order and names of phase3-issue-body.md:180-219,
    Slater block of phase4-issue-body.md:97-111.  Sub-module names are the ``setup`` attribute names
    of the spec (orbitals_up, orbitals_down, jastrow, en_cusp, ee_cusp, backflow)."""
    n_up: int
    n_down: int
    hidden_dims: Tuple[int, ...] = (32, 32)
    use_cusp: bool = True
    use_backflow: bool = True

    @nn.compact
    def __call__(self, r, r_nuclei, charges, spin_mask):
        n = r.shape[0]
        diff = r[:, None, :] - r[None, :, :]
        eye = jnp.eye(n)
        # Molecule.electron_electron_distances (molecule.py:114-126) with a differentiable zero diagonal
        r_ee = jnp.sqrt(jnp.sum(diff ** 2, axis=-1) + eye) * (1.0 - eye)
        r_en = jnp.linalg.norm(r[:, None, :] - r_nuclei[None, :, :], axis=-1)      # molecule.py:128-140
        r_bf = SPEC["BackflowTransform"](name="backflow")(r, r_ee) if (self.use_backflow and n > 1) else r
        sign, log_det = 1.0, 0.0
        for nm, lo, ns in (("orbitals_up", 0, self.n_up), ("orbitals_down", self.n_up, self.n_down)):
            if ns == 0:
                continue
            orb = SPEC["OrbitalSet"](ns, self.hidden_dims, name=nm)
            phi = jnp.stack([orb(r_bf[lo + i], r_nuclei, charges) for i in range(ns)])   # phase4:101-104
            s, la = SPEC["log_slater_determinant"](phi)                                   # phase2:38-50
            sign, log_det = sign * s, log_det + la
        log_j = SPEC["JastrowFactor"](name="jastrow")(r_ee, r_en)
        log_cusp = 0.0
        if self.use_cusp:
            log_cusp = log_cusp + SPEC["ElectronNuclearCusp"](name="en_cusp")(r_en, charges)
            log_cusp = log_cusp + SPEC["ElectronElectronCusp"](name="ee_cusp")(r_ee, spin_mask)
        return sign, log_det + log_j + log_cusp


class SpecWavefunction(Wavefunction):
    """The [SPEC] ansatz behind the reference's implemented Wavefunction ABC (base.py:38-76)."""

    def __init__(self, molecule: Molecule, hidden_dims, use_cusp, use_backflow):
        self.molecule = molecule
        self.mod = SpecAntisymmetric(molecule.spin[0], molecule.spin[1], tuple(hidden_dims), use_cusp, use_backflow)

    def _args(self):
        m = self.molecule
        return jnp.array(m.positions), jnp.array(m.charges), jnp.array(m.get_spin_mask())

    def sign_and_log(self, params, r):
        return self.mod.apply(params, r, *self._args())

    def __call__(self, params, r):
        return self.sign_and_log(params, r)[1]

    def init(self, key, r_sample):
        return self.mod.init(key, r_sample, *self._args())


# ----------------------------------------------------------------------------------------------
def np64(x):
    return np.asarray(torch.as_tensor(x).detach().to(torch.float64).numpy())


def cast_tree(t, dtype):
    return jax.tree.map(lambda x: x.to(dtype), t)


def tape_pop():
    t = list(jax.random.TAPE)
    jax.random.TAPE.clear()
    return t


def run_case(name, molecule, make_wf, W, step_size, seed, vmc_steps=0, is_spec=False):
    out, meta = {}, {"name": name, "W": W, "step_size": step_size, "seed": seed}
    jax.config.update("jax_enable_x64", True)
    wf = make_wf()
    key = jax.random.PRNGKey(seed)
    k_init, k_walk, k_burn, k_mh, k_smp, k_vmc = jax.random.split(key, 6)
    n_el = molecule.n_electrons
    params = wf.init(k_init, jnp.zeros((n_el, 3)))
    if is_spec:
        # the spec initialises Jastrow/cusp scalars to one and biases to zero; perturb every leaf a little
        # so no term of the gradient is checked at a symmetric point
        g = np.random.default_rng(seed + 99)
        amp = 0.05
    else:
        g = np.random.default_rng(seed + 99)
        amp = 0.02
    params = jax.tree.map(lambda x: (x + torch.as_tensor(amp * g.standard_normal(tuple(x.shape))))
                          .to(torch.float32).to(torch.float64), params)     # float32-representable
    paths = jax.tree.paths(params)
    meta["leaves"] = [[p, list(v.shape)] for p, v in paths]
    theta = np64(jax.tree.ravel(params))
    assert np.array_equal(theta, theta.astype(np.float32).astype(np.float64))
    out["theta"] = theta.astype(np.float32)
    meta["P"] = int(theta.size)

    # walkers: the reference's own initialisation + its own sampler as burn-in  (vmc.py:161-215)
    opt = VMCOptimizer(n_chains=W, step_size=step_size)
    st = opt.init(k_walk, wf, params, molecule)
    log_prob_fn = lambda x: wf.log_prob(params, x)   # noqa: E731  (vmc.py:238-239)
    sampler = MetropolisSampler(step_size=step_size)
    _, sst, _ = sampler.sample(k_burn, log_prob_fn, st.sampler_state, n_samples=1, n_burn=40)
    r = sst.positions.to(torch.float32).to(torch.float64)          # float32-representable inputs
    tape_pop()
    out["r"] = np64(r)

    def hot_path(prefix, params, r, dt):
        o = {}
        lp = jax.vmap(lambda x: wf(params, x))(r)
        o["log_psi"] = np64(lp)
        if is_spec:
            o["sign"] = np64(jax.vmap(lambda x: wf.sign_and_log(params, x)[0])(r))
        o["T"] = np64(jax.vmap(lambda x: kinetic_energy(lambda q: wf(params, q), x))(r))
        o["V_en"] = np64(jax.vmap(lambda x: electron_nuclear_potential(x, molecule))(r))
        o["V_ee"] = np64(jax.vmap(lambda x: electron_electron_potential(x))(r))
        o["E_L"] = np64(local_energy_batched(wf, params, r, molecule))
        # one MH step from (r, logp = 2 log|psi|)
        from nqmc.sampler.metropolis import SamplerState
        s0 = SamplerState(positions=r, log_prob=2.0 * lp, n_accepted=jnp.zeros(r.shape[0]), n_steps=0)
        tape_pop()
        s1 = sampler.step(k_mh, s0, lambda x: wf.log_prob(params, x))
        tape = tape_pop()
        assert [k for k, _ in tape] == ["normal", "uniform"]
        o["mh_normals"], o["mh_uniforms"] = tape[0][1], tape[1][1]
        prop = r + step_size * torch.as_tensor(tape[0][1], dtype=dt)
        o["mh_ratio"] = np64(jax.vmap(lambda x: wf.log_prob(params, x))(prop) - 2.0 * lp)
        o["mh_positions_out"], o["mh_logp_out"] = np64(s1.positions), np64(s1.log_prob)
        o["mh_accept"] = np64(s1.n_accepted).astype(np.int8)
        # sample(): burn-in, thinning, acceptance bookkeeping
        n_burn, n_samples, n_thin = 3, 2, 2
        smp, sf, acc = sampler.sample(k_smp, lambda x: wf.log_prob(params, x), s0, n_samples=n_samples,
                                      n_burn=n_burn, n_thin=n_thin)
        tape = tape_pop()
        o["smp_normals"] = np.stack([a for k, a in tape if k == "normal"])
        o["smp_uniforms"] = np.stack([a for k, a in tape if k == "uniform"])
        o["smp_samples"], o["smp_acc_rate"] = np64(smp), np64(acc)
        o["smp_final_logp"], o["smp_final_n_accepted"] = np64(sf.log_prob), np64(sf.n_accepted)
        # gradient estimator on those samples
        grad, mean_e, std_e = compute_gradient(wf, params, smp, molecule)
        o["grad"], o["grad_mean_energy"], o["grad_energy_std"] = np64(jax.tree.ravel(grad)), np64(mean_e), np64(std_e)
        em, es = estimate_energy(wf, params, smp, molecule)
        o["est_mean"], o["est_std"] = np64(em), np64(es)
        flat = smp.reshape(-1, *smp.shape[2:])
        if theta.size <= 50000:
            O = jax.vmap(lambda x: jax.tree.ravel(jax.grad(lambda p: wf(p, x))(params)))(flat)
            o["mean_O"] = np64(O.mean(0))
        return {prefix + k: v for k, v in o.items()}, dict(n_burn=n_burn, n_samples=n_samples, n_thin=n_thin)

    o64, smeta = hot_path("", params, r, torch.float64)
    out.update(o64)
    meta.update(smeta)
    meta["V_nn"] = float(molecule.nuclear_repulsion_energy())

    # ---- VMCOptimizer.step trajectory (vmc.py:217-282), x64
    if vmc_steps:
        vopt = VMCOptimizer(learning_rate=1e-2, n_samples=2, n_chains=W, n_burn=4, step_size=step_size, clip_grad=1.0)
        st = vopt.init(k_vmc, wf, params, molecule)
        tape = tape_pop()
        out["vmc_init_normals"] = tape[0][1]
        out["vmc_init_positions"], out["vmc_init_logp"] = np64(st.sampler_state.positions), np64(st.sampler_state.log_prob)
        thetas, mets, nrm, uni = [], [], [], []
        kk = k_vmc
        for i in range(vmc_steps):
            kk, sk = jax.random.split(kk)
            st, m = vopt.step(sk, st, wf, molecule)
            tape = tape_pop()
            nrm.append(np.stack([a for k, a in tape if k == "normal"]))
            uni.append(np.stack([a for k, a in tape if k == "uniform"]))
            thetas.append(np64(jax.tree.ravel(st.params)))
            mets.append([m["energy"], m["energy_std"], m["acceptance_rate"], m["grad_norm"]])
        out["vmc_theta"], out["vmc_metrics"] = np.stack(thetas), np.asarray(mets)
        for i in range(vmc_steps):
            out[f"vmc_normals_{i}"], out[f"vmc_uniforms_{i}"] = nrm[i], uni[i]
        out["vmc_final_positions"] = np64(st.sampler_state.positions)
        meta["vmc"] = dict(steps=vmc_steps, learning_rate=1e-2, n_samples=2, n_burn=4, clip_grad=1.0,
                           metric_keys=["energy", "energy_std", "acceptance_rate", "grad_norm"])

    # ---- the same hot path in the reference's shipped precision (float32)
    jax.config.update("jax_enable_x64", False)
    wf32 = make_wf()
    o32, _ = hot_path("f32_", cast_tree(params, torch.float32), r.to(torch.float32), torch.float32)
    for k in ("f32_log_psi", "f32_T", "f32_E_L", "f32_mh_accept", "f32_mh_ratio", "f32_grad", "f32_grad_mean_energy"):
        if k != "f32_grad" or theta.size <= 50000:
            out[k] = o32[k]
    del wf32
    jax.config.update("jax_enable_x64", True)
    return out, meta


def geometry(molecule):
    return dict(positions=np.asarray(molecule.positions).tolist(), charges=np.asarray(molecule.charges).tolist(),
                spin=list(molecule.spin))


CASES = {
    # name: (molecule, wavefunction factory, kwargs, description)
    "simple_hatom": (hydrogen_atom, lambda m: SimpleWavefunction(nuclear_positions=m.positions),
                     dict(W=48, step_size=0.5, seed=1), dict(kind="simple", hidden=[64, 64], activation="tanh")),
    "simple_h2": (hydrogen_molecule, lambda m: SimpleWavefunction(nuclear_positions=m.positions),
                  dict(W=48, step_size=0.2, seed=2, vmc_steps=3), dict(kind="simple", hidden=[64, 64], activation="tanh")),
    "simple_h2_silu3": (hydrogen_molecule,
                        lambda m: SimpleWavefunction(hidden_dims=(32, 16, 8), activation="silu", envelope_decay=1.3,
                                                     nuclear_positions=m.positions),
                        dict(W=48, step_size=0.2, seed=3, vmc_steps=2),
                        dict(kind="simple", hidden=[32, 16, 8], activation="silu", envelope_decay=1.3)),
    "simple_h4": (lambda: hydrogen_chain(4, 1.8), lambda m: SimpleWavefunction(nuclear_positions=m.positions),
                  dict(W=48, step_size=0.3, seed=4), dict(kind="simple", hidden=[64, 64], activation="tanh")),
    "spec_h2_h32": (hydrogen_molecule, lambda m: SpecWavefunction(m, (32, 32), False, False),
                    dict(W=32, step_size=0.2, seed=5, is_spec=True, vmc_steps=2),
                    dict(kind="antisymmetric", hidden=[32, 32], use_cusp=False, use_backflow=False)),
    "spec_h4_h64": (lambda: hydrogen_chain(4, 1.8), lambda m: SpecWavefunction(m, (64, 64), False, False),
                    dict(W=32, step_size=0.5, seed=6, is_spec=True),
                    dict(kind="antisymmetric", hidden=[64, 64], use_cusp=False, use_backflow=False)),
    "spec_h4_h64_cusp": (lambda: hydrogen_chain(4, 1.8), lambda m: SpecWavefunction(m, (64, 64), True, False),
                         dict(W=32, step_size=0.5, seed=7, is_spec=True),
                         dict(kind="antisymmetric", hidden=[64, 64], use_cusp=True, use_backflow=False)),
    "spec_h4_h64_bf_cusp": (lambda: hydrogen_chain(4, 1.8), lambda m: SpecWavefunction(m, (64, 64), True, True),
                            dict(W=24, step_size=0.5, seed=8, is_spec=True),
                            dict(kind="antisymmetric", hidden=[64, 64], use_cusp=True, use_backflow=True)),
    "spec_h6_h32_bf_cusp": (lambda: hydrogen_chain(6, 1.8), lambda m: SpecWavefunction(m, (32, 32), True, True),
                            dict(W=16, step_size=0.3, seed=9, is_spec=True),
                            dict(kind="antisymmetric", hidden=[32, 32], use_cusp=True, use_backflow=True)),
    "spec_h4_h128": (lambda: hydrogen_chain(4, 1.8), lambda m: SpecWavefunction(m, (128, 128), False, False),
                     dict(W=16, step_size=0.5, seed=10, is_spec=True),
                     dict(kind="antisymmetric", hidden=[128, 128], use_cusp=False, use_backflow=False)),
}


def sha(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def main(argv):
    names = argv or list(CASES)
    man_path = os.path.join(HERE, "ref_hotpath_manifest.json")
    manifest = json.load(open(man_path)) if os.path.exists(man_path) else {"cases": {}}
    manifest["generator"] = "tests/golden/make_ref_hotpath_golden.py"
    manifest["reference_files"] = {p: sha(os.path.join(REF, p)) for p in [
        "src/nqmc/hamiltonian/molecular.py", "src/nqmc/wavefunction/simple.py", "src/nqmc/wavefunction/base.py",
        "src/nqmc/sampler/metropolis.py", "src/nqmc/optimizer/vmc.py", "src/nqmc/systems/molecule.py",
        ".github/phase2-issue-body.md", ".github/phase3-issue-body.md"]}
    manifest["spec_blocks"] = [list(b) for b in SPEC_BLOCKS]
    manifest["substrate"] = f"oracle/_ref_shim over torch {torch.__version__} (jax/flax/optax not installable)"
    for name in names:
        mk_mol, mk_wf, kw, desc = CASES[name]
        mol = mk_mol()
        out, meta = run_case(name, mol, lambda: mk_wf(mol), **kw)
        meta.update(desc)
        meta["molecule"] = geometry(mol)
        np.savez_compressed(os.path.join(HERE, f"ref_hotpath_{name}.npz"), **out)
        manifest["cases"][name] = meta
        print(f"{name}: P={meta['P']} W={meta['W']} <E_L>={out['E_L'].mean():+.5f} acc={out['mh_accept'].mean():.2f} "
              f"|f32-f64| E_L med {np.median(np.abs(out['f32_E_L'] - out['E_L'])):.1e} "
              f"size={os.path.getsize(os.path.join(HERE, f'ref_hotpath_{name}.npz')) / 1e3:.0f} kB")
    json.dump(manifest, open(man_path, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main(sys.argv[1:])
