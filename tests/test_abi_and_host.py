"""CPU tests: the C-ABI library loads and exports every symbol include/nqmc_b200.h declares (no compute calls),
fails loudly without a GPU, and the host-side logic (parameter flattening, keys, Adam, sharding) is right."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

import nqmc_b200
from nqmc_b200 import _native, parallel, params as P, random as nqrandom
from nqmc_b200.optimizer import vmc
from nqmc_b200.systems import hydrogen_chain, hydrogen_molecule
from nqmc_b200.wavefunction import AntisymmetricWavefunction, SimpleWavefunction
from oracle.layout import KIND_SIMPLE, OracleNet, hydrogen_chain_system, leaf_table, param_count

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "nqmc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(nqmc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = _native.load_library()
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nqmc_b200.h but not exported"
    assert sorted(_native.EXPORTS) == declared
    assert lib.nqmc_abi_version() == 3


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _native.load_library()
    R = (ctypes.c_double * 6)(0, 0, 0, 1.4, 0, 0)
    Z = (ctypes.c_double * 2)(1, 1)
    s = _native._System(2, 2, 1, 0, R, Z, 0.714)
    n = _native._Net(1, 32, 32, 0, 0, 0, 1.0, 0.5, 0.25)
    h = ctypes.c_void_p()
    rc = lib.nqmc_create(ctypes.byref(s), ctypes.byref(n), 0, ctypes.byref(h))
    assert rc == -2 and not h.value
    assert b"no CPU fallback" in lib.nqmc_last_error(None)
    wf = SimpleWavefunction(nuclear_positions=hydrogen_molecule().positions)
    p = wf.init(nqrandom.PRNGKey(0), np.zeros((2, 3)))
    with pytest.raises(_native.NativeError):
        wf(p, np.zeros((2, 3)))
    from nqmc_b200.sampler import MetropolisSampler
    with pytest.raises(TypeError):
        MetropolisSampler().init_state(nqrandom.PRNGKey(0), lambda r: -np.sum(r * r), 4, 2)


def test_argument_validation_needs_no_gpu():
    lib = _native.load_library()
    h = ctypes.c_void_p()
    assert lib.nqmc_create(None, None, 0, ctypes.byref(h)) == -1
    R = (ctypes.c_double * 3)(0, 0, 0)
    Z = (ctypes.c_double * 1)(1)
    bad = _native._System(1, 40, 1, 0, R, Z, 0.0)
    n = _native._Net(1, 64, 64, 0, 0, 0, 1.0, 0.5, 0.25)
    assert lib.nqmc_create(ctypes.byref(bad), ctypes.byref(n), 0, ctypes.byref(h)) == -1
    ok = _native._System(1, 1, 1, 0, R, Z, 0.0)
    n2 = _native._Net(1, 48, 48, 0, 0, 0, 1.0, 0.5, 0.25)
    assert lib.nqmc_create(ctypes.byref(ok), ctypes.byref(n2), 0, ctypes.byref(h)) == -3      # width not 32/64/128
    assert lib.nqmc_param_count(None) == -1


@pytest.mark.parametrize("n_atoms,hidden,cusp,bf", [(2, 32, False, False), (4, 64, False, False), (6, 128, True, True)])
def test_param_tree_flattening_matches_oracle_layout(n_atoms, hidden, cusp, bf):
    mol = hydrogen_chain(n_atoms)
    wf = AntisymmetricWavefunction(n_atoms // 2, n_atoms // 2, (hidden, hidden), use_cusp=cusp, use_backflow=bf, molecule=mol)
    tree = wf.init(nqrandom.PRNGKey(1))
    s = hydrogen_chain_system(n_atoms)
    net = OracleNet(hidden=(hidden, hidden), use_cusp=cusp, use_backflow=bf)
    ref = leaf_table(s, net)
    ours = list(P.tree_leaves_with_path(tree))
    assert [p for p, _ in ours] == [p for p, _, _ in ref]
    assert [tuple(v.shape) for _, v in ours] == [sh for _, sh, _ in ref]
    theta = P.ravel(tree)
    assert theta.shape == (param_count(s, net),) and theta.dtype == np.float32
    back = P.unravel(theta, tree)
    for (pa, a), (pb, b) in zip(P.tree_leaves_with_path(tree), P.tree_leaves_with_path(back)):
        assert pa == pb and np.array_equal(a, b)
    j = tree["params"]["jastrow"]
    assert all(float(j[k]) == 1.0 for k in ("a_ee", "b_ee", "a_en", "b_en"))


def test_simple_wavefunction_tree():
    wf = SimpleWavefunction(hidden_dims=(64, 64), nuclear_positions=hydrogen_molecule().positions)
    tree = wf.init(nqrandom.PRNGKey(42), np.zeros((2, 3)))
    assert P.ravel(tree).size == 4673 == wf.n_params
    s = hydrogen_chain_system(2, 1.4)
    assert [p for p, _ in P.tree_leaves_with_path(tree)] == [p for p, _, _ in leaf_table(s, OracleNet(kind=KIND_SIMPLE, hidden=(64, 64)))]
    k = tree["params"]["Dense_0"]["kernel"]
    assert abs(k.std() - 1 / np.sqrt(6)) < 0.08 and np.all(tree["params"]["Dense_0"]["bias"] == 0)
    with pytest.raises(ValueError):
        SimpleWavefunction(activation="relu")


def test_keys():
    k = nqrandom.PRNGKey(42)
    a, b = nqrandom.split(k)
    assert a.dtype == np.uint32 and not np.array_equal(a, b)
    assert np.array_equal(nqrandom.split(k)[0], a)
    x = nqrandom.normal(a, (20000,))
    assert abs(x.mean()) < 0.03 and abs(x.std() - 1) < 0.03
    u = nqrandom.uniform(b, (1000,))
    assert 0 <= u.min() and u.max() < 1


def test_adam_clip_matches_optax_formulas():
    rng = np.random.default_rng(0)
    p = {"a": rng.standard_normal(5).astype(np.float32)}
    st = vmc._opt_init(p)
    g = (10 * rng.standard_normal(5)).astype(np.float32)
    upd, st2 = vmc._opt_update(g, st, lr=1e-2, clip=1.0)
    gc = g / np.linalg.norm(g)                                  # clip_by_global_norm(1.0)
    mu, nu = 0.1 * gc, 0.001 * gc * gc
    ref = -1e-2 * (mu / 0.1) / (np.sqrt(nu / 0.001) + 1e-8)
    assert np.allclose(upd, ref, rtol=1e-5)
    small = (1e-3 * rng.standard_normal(5)).astype(np.float32)  # below the clip threshold: untouched
    upd2, _ = vmc._opt_update(small, st, lr=1e-2, clip=1.0)
    assert np.allclose(upd2, -1e-2 * small / (np.abs(small) + 1e-8), rtol=1e-4)
    assert st2["count"] == 1


def test_shard_range_partitions_walkers():
    for n, ws in ((65536, 8), (10, 3), (7, 8), (262144, 8)):
        spans = [parallel.shard_range(n, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_combine_gradient_equals_reference_estimator():
    from oracle.nqmc_oracle import Oracle
    from oracle.layout import init_theta
    s = hydrogen_chain_system(4)
    net = OracleNet(hidden=(8, 8))
    th = init_theta(s, net, 1)
    r = s.R[np.arange(4) % 4][None] + 0.5 * np.random.default_rng(0).standard_normal((50, 4, 3))
    g, mean, std, e, O = Oracle(s, net).energy_and_gradient(th, r)
    hdr = np.array([len(e), e.sum(), (e * e).sum(), 0, 0, 0, 0, 0.0])
    gsum = ((e - e.mean())[:, None] * O).sum(0)
    g2, m2, s2 = parallel.combine_gradient(hdr, gsum)
    assert np.allclose(g2, g, rtol=1e-12) and np.isclose(m2, mean) and np.isclose(s2, std, rtol=1e-8)


def test_checkpoint_roundtrip_host(tmp_path):
    from nqmc_b200.checkpoint import load_checkpoint, save_checkpoint
    from nqmc_b200.optimizer.vmc import VMCState
    from nqmc_b200.sampler import SamplerState
    mol = hydrogen_chain(4)
    wf = AntisymmetricWavefunction(2, 2, (32, 32), use_cusp=True, use_backflow=True, molecule=mol)
    params = wf.init(nqrandom.PRNGKey(1))
    rng = np.random.default_rng(0)
    n = P.ravel(params).size
    opt = {"count": 7, "mu": rng.standard_normal(n).astype(np.float32), "nu": rng.random(n).astype(np.float32)}
    ss = SamplerState(rng.standard_normal((16, 4, 3)).astype(np.float32), rng.standard_normal(16).astype(np.float32),
                      rng.integers(0, 5, 16).astype(np.float32), 11)
    st = VMCState(params, opt, ss, 7)
    save_checkpoint(str(tmp_path / "c.npz"), st, energy=-2.01, key=nqrandom.PRNGKey(9))
    st2, e, key = load_checkpoint(str(tmp_path / "c.npz"), wf.init(nqrandom.PRNGKey(2)))
    assert e == -2.01 and st2.step == 7 and st2.opt_state["count"] == 7 and np.array_equal(key, nqrandom.PRNGKey(9))
    assert np.array_equal(P.ravel(st2.params), P.ravel(params))
    assert np.array_equal(st2.opt_state["mu"], opt["mu"]) and np.array_equal(st2.opt_state["nu"], opt["nu"])
    assert np.array_equal(st2.sampler_state.positions, ss.positions) and st2.sampler_state.n_steps == 11


def test_item_iterator_matches_div_mod():
    """The persistent orbital kernels advance (walker block, electron) incrementally instead of dividing per tile
    (csrc/nqmc_internal.cuh, ItemIter); the same struct, run on the host through the C ABI, against item // ns, item % ns."""
    import ctypes
    lib = _native.load_library()
    rng = np.random.default_rng(0)
    cases = [(0, 1, 1), (0, 37, 2), (36, 37, 2), (5, 148, 5), (147, 148, 3), (3, 7, 10), (0, 29, 4), (11, 14, 5)]
    cases += [(int(rng.integers(0, 200)), int(rng.integers(1, 200)), int(rng.integers(1, 11))) for _ in range(40)]
    n = 5000
    for first, stride, ns in cases:
        wb = np.empty(n, np.int32); e = np.empty(n, np.int32)
        rc = lib.nqmc_debug_item_iter(first, stride, ns, n, ctypes.c_void_p(wb.ctypes.data), ctypes.c_void_p(e.ctypes.data))
        assert rc == 0
        items = first + stride * np.arange(n, dtype=np.int64)
        assert np.array_equal(wb, items // ns), (first, stride, ns)
        assert np.array_equal(e, items % ns), (first, stride, ns)
    bad = np.empty(1, np.int32)
    assert lib.nqmc_debug_item_iter(0, 0, 2, 1, ctypes.c_void_p(bad.ctypes.data), ctypes.c_void_p(bad.ctypes.data)) != 0
