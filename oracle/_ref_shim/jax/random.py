"""ORACLE / TEST INFRASTRUCTURE ONLY -- stand-in for `jax.random`.

NOT threefry.  A key is two 32-bit words (int64 tensor of shape (2,)); `split` and the samplers
run a NumPy Philox generator keyed by those words, so runs are deterministic.  Every draw is
appended to TAPE as (kind, shape, float64 ndarray): the fixture generator exports the proposal /
uniform streams from it, which is how parity "on identical streams" is set up (the JAX version
is unpinned, so reproducing threefry bit for bit would pin nothing).
"""
from __future__ import annotations

import numpy as _np
import torch

from ._config import float_dtype, to_torch_dtype

TAPE = []          # [(kind, ndarray)] in draw order; clear with TAPE.clear()
PRNGKey = None     # set below (a function; the reference also uses it as a type annotation)


def _gen(key, salt):
    k = [int(v) & 0xFFFFFFFF for v in torch.as_tensor(key).reshape(-1)[-2:].tolist()]
    return _np.random.Generator(_np.random.Philox(key=[(k[0] << 32) | k[1], salt]))


def _prngkey(seed):
    return torch.tensor([0, int(seed) & 0xFFFFFFFF], dtype=torch.int64)


PRNGKey = _prngkey
key = _prngkey


def split(key, num=2):
    w = _gen(key, 0x5EED).integers(0, 2 ** 32, size=(num, 2), dtype=_np.uint64)
    return torch.as_tensor(w.astype(_np.int64))


def normal(key, shape=(), dtype=None):
    x = _gen(key, 1).standard_normal(tuple(shape), dtype=_np.float32).astype(_np.float64)   # float32-representable
    dt = to_torch_dtype(dtype) if dtype is not None else float_dtype()
    x = torch.as_tensor(x, dtype=dt)           # rounded to the working precision first,
    TAPE.append(("normal", x.to(torch.float64).numpy().copy()))   # so the tape holds what was used
    return x


def uniform(key, shape=(), dtype=None, minval=0.0, maxval=1.0):
    """U[minval, maxval); float32-representable values so both precision modes and the CUDA path
    consume identical numbers."""
    u = _gen(key, 2).random(tuple(shape), dtype=_np.float32).astype(_np.float64)
    x = minval + (maxval - minval) * u
    dt = to_torch_dtype(dtype) if dtype is not None else float_dtype()
    x = torch.as_tensor(x, dtype=dt)
    TAPE.append(("uniform", x.to(torch.float64).numpy().copy()))
    return x
