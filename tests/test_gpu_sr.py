"""Stochastic Reconfiguration on the device: TMA + tcgen05 3xTF32 SYRK, CG solve.
Oracle: oracle/sr.py (fp64 restatement of .github/phase6-issue-body.md:150-181)."""
import numpy as np
import pytest

from _util import make_case, make_context
from oracle.sr import sr_matrices, sr_update

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    s, net, theta, _ = make_case(n_atoms=2, hidden=32, W=8, burn=0)
    c = make_context(s, net)
    c.set_params(theta)
    return c


def _fp32_yardstick(A, B):
    """error of the same product evaluated in plain fp32 (NumPy) against fp64: the bar for 'fp32-grade'."""
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    f32 = (A @ B.T).astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    return ref, scale, float(np.max(np.abs(f32 - ref) / scale))


@pytest.mark.parametrize("M,N,K", [(128, 256, 32), (128, 256, 4096), (200, 72, 1000), (5, 3, 8), (384, 512, 352), (1, 1, 4),
                                   (300, 300, 20000)])
def test_gram_matches_fp64(ctx, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.standard_normal((M, K)).astype(np.float32) * rng.uniform(0.1, 10, (M, 1)).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    C = ctx.gram(ctx.upload(A), ctx.upload(B), alpha=0.5, diag=0.25).numpy(ctx.stream).astype(np.float64)
    ref, scale, e32 = _fp32_yardstick(A, B)
    ref = 0.5 * ref + 0.25 * np.eye(M, N)
    err = float(np.max(np.abs(C - ref) / (0.5 * scale + 1e-30)))
    # 3xTF32 drops the lo*lo term and the tensor core adds into its fp32 accumulator with truncation (the kernel cuts
    # the chain every 256 samples): measured ~1e-6 of sum|a||b| for mixed signs, up to 3e-6 for same-sign sums
    assert err <= 4e-6, (err, e32)


@pytest.mark.parametrize("P,n", [(96, 512), (300, 4096), (1000, 2048), (130, 36)])
def test_symmetric_gram_is_symmetric_and_right(ctx, P, n):
    rng = np.random.default_rng(P + n)
    A = rng.standard_normal((P, n)).astype(np.float32)
    C = ctx.gram(ctx.upload(A), alpha=1.0 / n, diag=1e-3).numpy(ctx.stream)
    assert np.array_equal(C, C.T)                      # mirrored element by element: exactly symmetric
    ref, scale, e32 = _fp32_yardstick(A, A)
    ref = ref / n + 1e-3 * np.eye(P)
    assert np.max(np.abs(C - ref) / (scale / n)) <= 4e-6    # diagonal = same-sign sums: the worst case of the truncating accumulator


def _sr_problem(P, n, seed, corr=0.3):
    rng = np.random.default_rng(seed)
    mix = np.eye(P) + corr * rng.standard_normal((P, P)) / np.sqrt(P)
    g = (rng.standard_normal((n, P)) @ mix + rng.standard_normal(P) * 2.0).astype(np.float32)   # non-zero means
    e = (-1.1 + 0.3 * rng.standard_normal(n) + 0.05 * g[:, 0]).astype(np.float32)
    return g, e


@pytest.mark.parametrize("P,n", [(64, 1024), (500, 4096), (1200, 8192)])
def test_sr_update_matches_oracle(ctx, P, n):
    g, e = _sr_problem(P, n, P)
    Ot = ctx.upload(np.ascontiguousarray(g.T))
    delta, info, S = ctx.sr_update(Ot, ctx.upload(e), reg=1e-3, lr=0.05, max_iter=400, tol=1e-8)
    delta, info, S = delta.numpy(ctx.stream), info.numpy(ctx.stream), S.numpy(ctx.stream)
    S_ref, f_ref = sr_matrices(g, e, 1e-3)
    assert np.max(np.abs(S - S_ref)) <= 5e-6 * np.max(np.abs(S_ref))   # diagonal: same-sign sums, truncating accumulator
    ref = sr_update(g, e, learning_rate=0.05, regularization=1e-3)
    assert info[2] == 1.0 and info[1] <= 1e-8, info
    # the solve amplifies the fp32 error of S by its condition number (<~ 1e4 here): 1e-4 of |delta|
    assert np.linalg.norm(delta - ref) <= 1e-4 * np.linalg.norm(ref), (np.linalg.norm(delta - ref), np.linalg.norm(ref))
    # the centred O is left in Ot (documented in-place behaviour)
    oc = Ot.numpy(ctx.stream)
    assert np.allclose(oc, (g - g.mean(0)).T, atol=2e-5)


def test_sr_drops_samples_with_nonfinite_energy(ctx):
    g, e = _sr_problem(80, 512, 5)
    e_bad = e.copy()
    e_bad[[3, 77, 400]] = [np.nan, np.inf, -np.inf]
    keep = np.isfinite(e_bad)
    delta, info, _ = ctx.sr_update(ctx.upload(np.ascontiguousarray(g.T)), ctx.upload(e_bad), reg=1e-3, lr=1.0, max_iter=300, tol=1e-9)
    ref = sr_update(g[keep], e_bad[keep], learning_rate=1.0, regularization=1e-3)
    d = delta.numpy(ctx.stream)
    assert np.linalg.norm(d - ref) <= 1e-4 * np.linalg.norm(ref)


def test_host_class_mirrors_the_spec_signature(ctx):
    from nqmc_b200.optimizer import StochasticReconfiguration
    g, e = _sr_problem(48, 333, 9)          # n not a multiple of 4: padded on upload
    sr = StochasticReconfiguration(learning_rate=1e-2, regularization=1e-4, context=ctx)
    d = sr.compute_update(g, e)
    ref = sr_update(g, e, 1e-2, 1e-4)
    assert d.shape == (48,) and np.linalg.norm(d - ref) <= 2e-4 * np.linalg.norm(ref)
    assert sr.last_info["converged"] == 1.0
    with pytest.raises(ValueError):
        sr.compute_update(g, e[:-1])


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_sr_update_matches_reference_golden(ctx, case):
    """CUDA path vs the update the reference's own [SPEC] code block produced (tests/golden/make_ref_sr_golden.py)."""
    import os
    from nqmc_b200.optimizer import StochasticReconfiguration
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_sr_golden.npz"))
    lr, reg = g[f"{case}_lr_reg"]
    sr = StochasticReconfiguration(learning_rate=float(lr), regularization=float(reg), context=ctx, tol=1e-9)
    d = sr.compute_update(g[f"{case}_grads"], g[f"{case}_energies"])
    ref = g[f"{case}_update"]
    # fp32 S with condition number up to ~1/reg: 2e-4 of |delta| (fp32 tolerance of the path, north_star: fp32)
    assert np.linalg.norm(d - ref) <= 2e-4 * np.linalg.norm(ref), (np.linalg.norm(d - ref) / np.linalg.norm(ref), sr.last_info)
