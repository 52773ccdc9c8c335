"""ORACLE / TEST INFRASTRUCTURE ONLY -- stand-in for `jax.numpy` over torch (see ../README.md).

Only what the reference's hot-path files and the [SPEC] blocks call.  dtype rule as in JAX:
floating inputs become the default float (float32, or float64 with jax_enable_x64), python
scalars are weakly typed.
"""
from __future__ import annotations

import types

import numpy as _np
import torch

from ._config import _DT, float_dtype, to_torch_dtype

ndarray = torch.Tensor
float32 = _DT(torch.float32)
float64 = _DT(torch.float64)
int32 = _DT(torch.int32)
int64 = _DT(torch.int64)
bool_ = _DT(torch.bool)
pi = _np.pi
inf = _np.inf
newaxis = None


def _t(x):
    """Anything -> tensor under the JAX default-dtype rule."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, (list, tuple)) and len(x) and any(isinstance(e, torch.Tensor) for e in x):
        return torch.stack([_t(e) for e in x])
    a = _np.asarray(x)
    if a.dtype.kind == "f":
        return torch.as_tensor(a, dtype=float_dtype())
    if a.dtype.kind == "b":
        return torch.as_tensor(a, dtype=torch.bool)
    return torch.as_tensor(a)


def array(x, dtype=None):
    t = _t(x)
    return t.to(to_torch_dtype(dtype)) if dtype is not None else t


asarray = array


def _shape(s):
    return (s,) if isinstance(s, int) else tuple(s)


def zeros(shape, dtype=None):
    return torch.zeros(_shape(shape), dtype=to_torch_dtype(dtype) if dtype else float_dtype())


def ones(shape, dtype=None):
    return torch.ones(_shape(shape), dtype=to_torch_dtype(dtype) if dtype else float_dtype())


def zeros_like(x):
    return torch.zeros_like(_t(x))


def ones_like(x):
    return torch.ones_like(_t(x))


def eye(n, dtype=None):
    return torch.eye(n, dtype=to_torch_dtype(dtype) if dtype else float_dtype())


def arange(*a):
    return torch.arange(*a)


def triu(x, k=0):
    return torch.triu(_t(x), diagonal=k)


def tril(x, k=0):
    return torch.tril(_t(x), diagonal=k)


def _axis(kw, axis):
    if axis is not None:
        kw["dim"] = axis
    return kw


def sum(x, axis=None):  # noqa: A001
    return torch.sum(_t(x), **_axis({}, axis))


def mean(x, axis=None):
    return torch.mean(_t(x), **_axis({}, axis))


def std(x, axis=None, ddof=0):
    """jnp.std: population standard deviation (ddof=0) -- optimizer/vmc.py:77,114."""
    return torch.std(_t(x), correction=ddof, **_axis({}, axis))


def var(x, axis=None, ddof=0):
    return torch.var(_t(x), correction=ddof, **_axis({}, axis))


def sqrt(x):
    return torch.sqrt(_t(x) if isinstance(x, torch.Tensor) else torch.as_tensor(float(x), dtype=float_dtype()))


def exp(x):
    return torch.exp(_t(x))


def log(x):
    return torch.log(_t(x))


def abs(x):  # noqa: A001
    return torch.abs(_t(x))


def tanh(x):
    return torch.tanh(_t(x))


def sign(x):
    return torch.sign(_t(x))


def square(x):
    return _t(x) ** 2


def maximum(a, b):
    a, b = _t(a), _t(b)
    return torch.maximum(a, b.to(a.dtype) if b.ndim == 0 else b)


def minimum(a, b):
    a, b = _t(a), _t(b)
    return torch.minimum(a, b.to(a.dtype) if b.ndim == 0 else b)


def where(c, a, b):
    c = _t(c).to(torch.bool)
    a, b = _t(a), _t(b)
    if a.ndim == 0 and not isinstance(a, torch.Tensor):
        a = a.to(float_dtype())
    return torch.where(c, a, b)


def trace(x):
    return torch.trace(x) if x.ndim == 2 else torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)


def stack(xs, axis=0):
    return torch.stack([_t(x) for x in xs], dim=axis)


def concatenate(xs, axis=0):
    return torch.cat([_t(x) for x in xs], dim=axis)


def reshape(x, shape):
    return _t(x).reshape(shape)


def dot(a, b):
    return _t(a) @ _t(b)


def allclose(a, b, rtol=1e-5, atol=1e-8):
    return bool(torch.allclose(_t(a), _t(b).to(_t(a).dtype), rtol=rtol, atol=atol))


def isfinite(x):
    return torch.isfinite(_t(x))


def _norm(x, axis=None, ord=None):  # noqa: A002
    """jnp.linalg.norm (2-norm): sqrt(sum x^2) -- NaN gradient at the zero vector, like JAX."""
    x = _t(x)
    if axis is None:
        return torch.sqrt(torch.sum(x * x))
    return torch.sqrt(torch.sum(x * x, dim=axis))


def _det(M):
    """Laplace expansion along row 0 (n <= 6): see README, torch.linalg.slogdet differentiates
    wrongly under vmap(hessian(.)) on torch 2.11."""
    n = M.shape[-1]
    if n == 1:
        return M[..., 0, 0]
    if n == 2:
        return M[..., 0, 0] * M[..., 1, 1] - M[..., 0, 1] * M[..., 1, 0]
    out = 0.0
    cols = list(range(n))
    for j in range(n):
        minor = M[..., 1:, cols[:j] + cols[j + 1:]]
        out = out + ((-1.0) ** j) * M[..., 0, j] * _det(minor)
    return out


def _slogdet(M):
    """jnp.linalg.slogdet -> (sign, log|det|)."""
    M = _t(M)
    if M.shape[-1] > 6:
        raise NotImplementedError("shim slogdet: n <= 6")
    d = _det(M)
    return torch.sign(d), torch.log(torch.abs(d))


def einsum(spec, *ops):
    return torch.einsum(spec, *[_t(o) for o in ops])


linalg = types.SimpleNamespace(norm=_norm, slogdet=_slogdet, det=lambda M: _det(_t(M)),
                               solve=lambda A, b: torch.linalg.solve(_t(A), _t(b)),
                               inv=lambda M: torch.linalg.inv(_t(M)))
