"""ORACLE / TEST INFRASTRUCTURE ONLY -- stand-in for `flax.linen` (see ../README.md).

Provides what the reference's wavefunction/simple.py:28-70 and the [SPEC] modules
(.github/phase2-issue-body.md:78-152, phase3-issue-body.md:50-169) use:

* ``Module``: subclasses become dataclasses over their annotated fields plus ``name=``;
  ``@compact`` ``__call__``; ``init(key, *args) -> {'params': tree}``; ``apply(variables, *args)``.
  Sub-modules created inside a compact call are named ``<ClassName>_<i>`` in creation order unless
  ``name=`` is given -- flax's auto-naming rule, which fixes the parameter tree
  (``Dense_0/{kernel,bias}``, ...).
* ``Dense(features)``: ``y = x @ kernel + bias``, kernel ``(in, out)`` LeCun-normal (truncated at
  2 sigma, unit-variance corrected), bias zeros -- flax defaults.
* ``self.param(name, init_fn, shape)``, ``initializers.ones/zeros``, ``tanh``, ``silu``.
"""
from __future__ import annotations

import types

import numpy as _np
import torch

from jax._config import float_dtype

_STACK = []          # active module calls (innermost last)
_MODE = {"init": False, "rng": None}


def compact(f):
    return f


def tanh(x):
    return torch.tanh(x)


def silu(x):
    return x * torch.sigmoid(x)


swish = silu


def _ones(key, shape, dtype=None):
    return torch.ones(tuple(shape), dtype=float_dtype())


def _zeros(key, shape, dtype=None):
    return torch.zeros(tuple(shape), dtype=float_dtype())


def _lecun_normal(key, shape, dtype=None):
    fan_in = shape[0]
    rng = _MODE["rng"]
    n = int(_np.prod(shape))
    x = rng.standard_normal(4 * n + 16)
    x = x[_np.abs(x) <= 2.0][:n] / 0.87962566103423978
    w = (x / _np.sqrt(fan_in)).astype(_np.float32).astype(_np.float64)     # float32-representable
    return torch.as_tensor(w.reshape(shape), dtype=float_dtype())


initializers = types.SimpleNamespace(ones=_ones, zeros=_zeros, lecun_normal=lambda: _lecun_normal)


class Module:
    name = None

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        fields = []
        for klass in reversed(cls.__mro__):
            for f in getattr(klass, "__annotations__", {}):
                if f not in fields and f not in ("name", "parent"):
                    fields.append(f)
        cls._fields = tuple(fields)
        call = cls.__dict__.get("__call__")
        if call is not None:
            cls._user_call = call

            def wrapped(self, *a, **k):
                return self._run(a, k)
            cls.__call__ = wrapped

    def __init__(self, *args, name=None, parent=None, **kw):
        if len(args) > len(self._fields):
            raise TypeError(f"{type(self).__name__}: too many positional arguments")
        vals = dict(zip(self._fields, args))
        for k, v in kw.items():
            if k not in self._fields:
                raise TypeError(f"{type(self).__name__}: unexpected field {k!r}")
            vals[k] = v
        for f in self._fields:
            if f in vals:
                object.__setattr__(self, f, vals[f])
            elif not hasattr(type(self), f):
                raise TypeError(f"{type(self).__name__}: missing field {f!r}")
        self.name = name
        self._scope = None
        self._counts = {}
        self._bound_name = None

    # -- scope handling -------------------------------------------------------------------
    def _child_name(self, child):
        if child.name is not None:
            return child.name
        if child._bound_name is None:            # one auto name per instance
            cls = type(child).__name__
            i = self._counts.get(cls, 0)
            self._counts[cls] = i + 1
            child._bound_name = f"{cls}_{i}"
        return child._bound_name

    def _run(self, a, k, top_scope=None):
        if top_scope is not None:
            self._scope = top_scope
        else:
            if not _STACK:
                raise RuntimeError("call a Module through .init(...) or .apply(...)")
            parent = _STACK[-1]
            nm = parent._child_name(self)
            self._scope = parent._scope.setdefault(nm, {}) if _MODE["init"] else parent._scope[nm]
        self._counts = {}
        _STACK.append(self)
        try:
            return type(self)._user_call(self, *a, **k)
        finally:
            _STACK.pop()

    def param(self, name, init_fn, shape=(), *a):
        sc = self._scope
        if _MODE["init"] and name not in sc:
            sc[name] = init_fn(None, tuple(shape))
        return sc[name]

    def _top(self, tree, init, rng, a, k):
        saved = (list(_STACK), dict(_MODE))
        _STACK.clear()
        _MODE.update(init=init, rng=rng)
        try:
            return self._run(a, k, top_scope=tree)
        finally:
            _STACK.clear()
            _STACK.extend(saved[0])
            _MODE.update(saved[1])
            self._scope = None

    def init(self, key, *a, **k):
        from jax.random import _gen
        tree = {}
        self._top(tree, True, _gen(key, 7), a, k)
        return {"params": tree}

    def apply(self, variables, *a, **k):
        return self._top(variables["params"], False, None, a, k)


class Dense(Module):
    features: int
    use_bias: bool = True

    def __call__(self, x):
        kernel = self.param("kernel", _lecun_normal, (x.shape[-1], self.features))
        y = x @ kernel
        if self.use_bias:
            y = y + self.param("bias", _zeros, (self.features,))
        return y
