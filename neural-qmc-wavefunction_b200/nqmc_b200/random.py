"""Host-side PRNG keys.

The reference threads ``jax.random`` keys through its API (``key`` arguments of
``MetropolisSampler.step`` / ``VMCOptimizer.step``, src/nqmc/sampler/metropolis.py:99-115,
src/nqmc/optimizer/vmc.py:308-314).  JAX is not a dependency here; a key is a ``uint32[2]`` array
and the device generator is counter-based Philox4x32-10 keyed by it (csrc/nqmc_internal.cuh).
Streams are therefore reproducible but not bit-equal to any JAX version; parity
tests pass explicit normal / uniform arrays instead.
"""
from __future__ import annotations

import numpy as np

_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85


def _philox(ctr, key):
    c = [int(x) for x in ctr]
    k0, k1 = int(key[0]), int(key[1])
    for _ in range(10):
        p0, p1 = c[0] * _M0, c[2] * _M1
        c = [((p1 >> 32) ^ c[1] ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, ((p0 >> 32) ^ c[3] ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k0 = (k0 + _W0) & 0xFFFFFFFF
        k1 = (k1 + _W1) & 0xFFFFFFFF
    return c


def PRNGKey(seed: int) -> np.ndarray:
    seed = int(seed)
    return np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)


def split(key, num: int = 2) -> np.ndarray:
    """``num`` statistically independent child keys, shape (num, 2)."""
    key = np.asarray(key, dtype=np.uint32).reshape(2)
    out = np.empty((num, 2), dtype=np.uint32)
    for i in range(num):
        w = _philox([i & 0xFFFFFFFF, (i >> 32) & 0xFFFFFFFF, 0x73706C74, 0], key)   # domain tag 'splt'
        out[i] = (w[0], w[1])
    return out


def key_to_seed(key) -> int:
    key = np.asarray(key, dtype=np.uint32).reshape(2)
    return int(key[0]) | (int(key[1]) << 32)


def _generator(key) -> np.random.Generator:
    return np.random.Generator(np.random.Philox(key=key_to_seed(key)))


def normal(key, shape, dtype=np.float32) -> np.ndarray:
    return _generator(key).standard_normal(shape).astype(dtype)


def uniform(key, shape, dtype=np.float32) -> np.ndarray:
    return _generator(key).random(shape, dtype=np.float32).astype(dtype)
