"""ORACLE / TEST INFRASTRUCTURE ONLY -- stand-in for the `optax` import name (see ../README.md).

Third-party dependency of the reference (pyproject: optax>=0.1.7, unpinned, not vendored).  The
four calls at optimizer/vmc.py:180-184,256-261 are restated from optax's published definitions:

* clip_by_global_norm(c): g_norm = sqrt(sum_leaves sum g^2); g <- g if g_norm < c else g / g_norm * c
* adam(lr, b1=0.9, b2=0.999, eps=1e-8, eps_root=0): mu <- b1 mu + (1-b1) g; nu <- b2 nu + (1-b2) g^2;
  count <- count+1; mhat = mu / (1-b1^count); vhat = nu / (1-b2^count);
  u = -lr * mhat / (sqrt(vhat + eps_root) + eps)
* chain(t1, t2): states kept as a tuple, updates threaded left to right
* apply_updates(p, u) = p + u
"""
from __future__ import annotations

from typing import NamedTuple

import torch

import jax


class GradientTransformation(NamedTuple):
    init: callable
    update: callable


class EmptyState(NamedTuple):
    pass


class ScaleByAdamState(NamedTuple):
    count: torch.Tensor
    mu: dict
    nu: dict


def global_norm(updates):
    return torch.sqrt(sum(torch.sum(g * g) for g in jax.tree.leaves(updates)))


def clip_by_global_norm(max_norm):
    def init(params):
        return EmptyState()

    def update(updates, state, params=None):
        g_norm = global_norm(updates)
        trigger = g_norm < max_norm
        updates = jax.tree.map(lambda t: torch.where(trigger, t, (t / g_norm) * max_norm), updates)
        return updates, state
    return GradientTransformation(init, update)


def scale_by_adam(b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    def init(params):
        z = lambda p: torch.zeros_like(p)
        return ScaleByAdamState(torch.zeros((), dtype=torch.int32), jax.tree.map(z, params), jax.tree.map(z, params))

    def update(updates, state, params=None):
        mu = jax.tree.map(lambda g, m: (1 - b1) * g + b1 * m, updates, state.mu)
        nu = jax.tree.map(lambda g, v: (1 - b2) * (g * g) + b2 * v, updates, state.nu)
        count = state.count + 1
        c = float(count)
        mu_hat = jax.tree.map(lambda m: m / (1 - b1 ** c), mu)
        nu_hat = jax.tree.map(lambda v: v / (1 - b2 ** c), nu)
        out = jax.tree.map(lambda m, v: m / (torch.sqrt(v + eps_root) + eps), mu_hat, nu_hat)
        return out, ScaleByAdamState(count, mu, nu)
    return GradientTransformation(init, update)


def scale(step):
    return GradientTransformation(lambda p: EmptyState(),
                                  lambda u, s, params=None: (jax.tree.map(lambda g: step * g, u), s))


def chain(*ts):
    def init(params):
        return tuple(t.init(params) for t in ts)

    def update(updates, state, params=None):
        new = []
        for t, s in zip(ts, state):
            updates, s2 = t.update(updates, s, params)
            new.append(s2)
        return updates, tuple(new)
    return GradientTransformation(init, update)


def adam(learning_rate, b1=0.9, b2=0.999, eps=1e-8, eps_root=0.0):
    return chain(scale_by_adam(b1, b2, eps, eps_root), scale(-learning_rate))


def sgd(learning_rate):
    return scale(-learning_rate)


def apply_updates(params, updates):
    return jax.tree.map(lambda p, u: p + u, params, updates)
