"""ORACLE (test infrastructure only) -- Philox4x32-10 and the walker streams, in NumPy.

The reference draws its proposal normals and acceptance uniforms from JAX's threefry
(src/nqmc/sampler/metropolis.py:115-121,132).  JAX versions are unpinned and its stream
layout changed between releases, so bit-reproducing it is neither possible
nor required: parity uses externally supplied streams.  For throughput mode the device
generates its own counter-based stream; this file restates that generator so the integer
stream can be checked bit-exactly and the float transforms to 1 ulp-ish tolerance.

Algorithm: Philox4x32-10 (Salmon et al., SC'11), multipliers 0xD2511F53 / 0xCD9E8D57,
Weyl constants 0x9E3779B9 / 0xBB67AE85 -- the published constants.

Stream layout (must match csrc/nqmc_rng.cuh):
    key     = (seed_lo, seed_hi)
    counter = (walker_global_id, block_index, step_lo, step_hi)
    block 0                -> words[0] is the acceptance uniform
    block 1 + q, q>=0      -> 4 words -> 2 Box-Muller pairs -> normals 4q .. 4q+3
    uniform  u = (x >> 8) * 2^-24                  in [0,1)
    normal   (x0,x1): a = ((x0>>8)+1) * 2^-24 in (0,1];  b = (x1>>8) * 2^-24
             n0 = sqrt(-2 ln a) cos(2 pi b),  n1 = sqrt(-2 ln a) sin(2 pi b)
    normal index for electron i, axis d is 3*i+d.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(ctr: np.ndarray, key: np.ndarray) -> np.ndarray:
    """ctr (...,4) uint32, key (2,) uint32 -> (...,4) uint32."""
    c = [ctr[..., i].astype(np.uint32) for i in range(4)]
    k0, k1 = np.uint32(key[0]), np.uint32(key[1])
    for _ in range(10):
        p0 = c[0].astype(np.uint64) * M0
        p1 = c[2].astype(np.uint64) * M1
        hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), p0.astype(np.uint32)
        hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), p1.astype(np.uint32)
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = np.uint32((int(k0) + int(W0)) & 0xFFFFFFFF)
        k1 = np.uint32((int(k1) + int(W1)) & 0xFFFFFFFF)
    return np.stack(c, axis=-1)


def walker_streams(seed: int, step: int, walker_ids: np.ndarray, n_elec: int):
    """Returns (normals (W,n_elec,3) float32, uniforms (W,) float32, raw words) for one MH step."""
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    Wn = walker_ids.shape[0]
    nn = 3 * n_elec
    nblk = 1 + (nn + 3) // 4
    ctr = np.zeros((Wn, nblk, 4), dtype=np.uint32)
    ctr[..., 0] = walker_ids.astype(np.uint32)[:, None]
    ctr[..., 1] = np.arange(nblk, dtype=np.uint32)[None, :]
    ctr[..., 2] = np.uint32(step & 0xFFFFFFFF)
    ctr[..., 3] = np.uint32((step >> 32) & 0xFFFFFFFF)
    words = philox4x32_10(ctr, key)                       # (W,nblk,4)
    u = ((words[:, 0, 0] >> np.uint32(8)).astype(np.float64) * 2.0 ** -24).astype(np.float32)
    x = words[:, 1:, :].reshape(Wn, -1, 2)                # pairs
    a = ((x[..., 0] >> np.uint32(8)).astype(np.float64) + 1.0) * 2.0 ** -24
    b = (x[..., 1] >> np.uint32(8)).astype(np.float64) * 2.0 ** -24
    rad = np.sqrt(-2.0 * np.log(a))
    n = np.stack([rad * np.cos(2 * np.pi * b), rad * np.sin(2 * np.pi * b)], axis=-1).reshape(Wn, -1)
    normals = n[:, :nn].reshape(Wn, n_elec, 3).astype(np.float32)
    return normals, u, words
