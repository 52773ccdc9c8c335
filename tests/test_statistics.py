"""CPU: nqmc_b200.analysis against values produced by executing the reference's [SPEC] statistics code
(tests/golden/make_stats_golden.py), and the process-group plumbing without torch."""
import json
import multiprocessing as mp
import os

import numpy as np
import pytest

from nqmc_b200.analysis import blocking_analysis, compute_autocorrelation, effective_sample_size, reblocking_curve

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "stats_golden.json")))


def _series(c):
    rng = np.random.default_rng(c["seed"])
    n, rho = c["n"], c["rho"]
    x = np.empty(n)
    x[0] = rng.standard_normal()
    for i in range(1, n):
        x[i] = rho * x[i - 1] + np.sqrt(1 - rho * rho) * rng.standard_normal()
    return x - 1.17


@pytest.mark.parametrize("c", GOLD["cases"], ids=[f"rho{c['rho']}" for c in GOLD["cases"]])
def test_statistics_match_the_spec_code(c):
    x = _series(c)
    for nb in (20, 7):
        m, s = blocking_analysis(x, nb)
        assert np.isclose(m, c[f"blocking_{nb}"][0], rtol=1e-12) and np.isclose(s, c[f"blocking_{nb}"][1], rtol=1e-12)
    assert np.allclose(compute_autocorrelation(x, 8), c["acf_first8"], rtol=1e-10, atol=1e-14)
    assert np.isclose(effective_sample_size(x), c["ess"], rtol=1e-9)


def test_blocking_sees_autocorrelation():
    """correlated series: the blocked error bar exceeds the naive std/sqrt(n) by ~sqrt((1+rho)/(1-rho))."""
    c = [k for k in GOLD["cases"] if k["rho"] == 0.7][0]
    x = _series(c)
    naive = x.std() / np.sqrt(len(x))
    _, blocked = blocking_analysis(x, 20)
    assert 1.5 < blocked / naive < 3.5
    curve = reblocking_curve(x)
    assert curve[0][0] == 1 and abs(curve[0][1] - x.std(ddof=1) / np.sqrt(len(x))) < 1e-12
    assert max(s for _, s in curve[:8]) > 1.5 * curve[0][1]
    with pytest.raises(ValueError):
        blocking_analysis(x[:5], 20)


def _fg_worker(rank, world, root, q):
    import sys
    base = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [base, os.path.join(base, "neural-qmc-wavefunction_b200")]
    from nqmc_b200 import parallel
    parallel.set_group(parallel.FileGroup(rank, world, root, timeout=60))
    assert parallel.world() == (rank, world)
    got = parallel.group().exchange({"rank": rank, "blob": bytes([rank]) * 64})
    hdr = np.array([10.0 + rank, 1.0, 2.0, 3.0, 0, 0, 0, 0])
    parallel.all_reduce_sum_(hdr)
    lo, hi = parallel.shard_range(37, rank, world)
    parallel.group().barrier()
    q.put((rank, [g["rank"] for g in got], hdr.tolist(), (lo, hi)))


def test_file_group_world_3_without_torch(tmp_path):
    """FileGroup: the rendezvous nqmc_b200.parallel uses for the IPC-handle exchange when no torch.distributed exists."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_fg_worker, args=(r, 3, str(tmp_path / "rdv"), q)) for r in range(3)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    spans = []
    for rank, order, hdr, span in res:
        assert order == [0, 1, 2] and hdr[0] == 33.0 and hdr[1] == 3.0
        spans.append(span)
    assert spans[0][0] == 0 and spans[-1][1] == 37 and all(spans[i][1] == spans[i + 1][0] for i in range(2))


def test_product_package_imports_without_torch():
    import subprocess
    import sys
    base = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys; sys.path.insert(0, %r); import nqmc_b200; from nqmc_b200 import sampler, optimizer, sweep, parallel, "
            "checkpoint, analysis, hamiltonian, wavefunction; "
            "assert 'torch' not in sys.modules and 'jax' not in sys.modules; print('ok')"
            % os.path.join(base, "neural-qmc-wavefunction_b200"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.stdout.strip() == "ok", out.stderr[-800:]
