"""Blocking analysis, autocorrelation and effective sample size.

Signatures and definitions follow the reference's [SPEC] starter code, .github/phase5-issue-body.md:112-151
(``blocking_analysis(data, n_blocks=20) -> (mean, std_error)``, ``compute_autocorrelation(data, max_lag=100)``,
``effective_sample_size(data)``); tests/test_statistics.py pins them to values produced by executing that code block.
These three act on a 1-D host series (a few thousand numbers: one mean energy per iteration, or one chain).

``chain_blocking`` is the batched device version for the walker loop: ``e_l`` holds n_samples consecutive local
energies for each of W independent chains (the layout ``VMCOptimizer.step`` produces, flat index s * W + w); every
chain is cut into n_blocks blocks, and the error of the grand mean is std(block means)/sqrt(n_blocks * W).  With
n_blocks = n_samples it reduces to the reference's naive ``std/sqrt(n)`` (optimizer/vmc.py:114), which ignores the
autocorrelation of successive Metropolis samples and is therefore too small.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np


def blocking_analysis(data: np.ndarray, n_blocks: int = 20) -> Tuple[float, float]:
    """(mean of block means, std of block means / sqrt(n_blocks)) with block_size = len(data) // n_blocks
    (the tail that does not fill a block is dropped) -- phase5-issue-body.md:112-131."""
    data = np.asarray(data, dtype=np.float64)
    block_size = len(data) // n_blocks
    if block_size < 1:
        raise ValueError(f"need at least n_blocks={n_blocks} samples, got {len(data)}")
    b = data[:n_blocks * block_size].reshape(n_blocks, block_size).mean(axis=1)
    return float(b.mean()), float(b.std() / np.sqrt(n_blocks))


def compute_autocorrelation(data: np.ndarray, max_lag: int = 100) -> np.ndarray:
    """acf[lag] = mean((x[:n-lag] - mean)(x[lag:] - mean)) / var, lag = 0..max_lag-1 -- phase5-issue-body.md:134-145."""
    x = np.asarray(data, dtype=np.float64)
    n = len(x)
    d = x - x.mean()
    var = x.var()
    return np.array([np.mean(d[:n - lag] * d[lag:]) / var for lag in range(max_lag)])


def effective_sample_size(data: np.ndarray) -> float:
    """n / (1 + 2 sum_{lag>=1} acf[lag]) -- phase5-issue-body.md:148-151."""
    acf = compute_autocorrelation(data)
    return float(len(data) / (1.0 + 2.0 * np.sum(acf[1:])))


def reblocking_curve(data: np.ndarray):
    """Flyvbjerg-Petersen curve: [(block_size, std_error)] for block sizes 1, 2, 4, ...; the plateau is the error bar."""
    x = np.asarray(data, dtype=np.float64)
    out, size = [], 1
    while len(x) >= 2:
        out.append((size, float(x.std(ddof=1) / np.sqrt(len(x)))))
        x = 0.5 * (x[0:len(x) // 2 * 2:2] + x[1:len(x) // 2 * 2:2])
        size *= 2
    return out


def chain_blocking(ctx, e_l, n_samples: int, n_chains: int, n_blocks: int) -> Tuple[float, float]:
    """(mean, std_error) over W chains x n_samples energies on the device (nqmc_blocked_stats); ``e_l`` is a device
    buffer of n_samples * n_chains floats, sample-major.  Non-finite energies invalidate their block."""
    out = ctx.blocked_stats(e_l, n_samples, n_chains, n_blocks)
    cnt, s1, s2 = (float(v) for v in out)
    if cnt <= 0:
        raise FloatingPointError("no finite block")
    mean = s1 / cnt
    var = max(s2 / cnt - mean * mean, 0.0)
    return mean, float(np.sqrt(var) / np.sqrt(cnt))
