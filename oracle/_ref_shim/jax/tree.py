"""ORACLE / TEST INFRASTRUCTURE ONLY -- `jax.tree.map` / `jax.tree.leaves` for dict / tuple / list
pytrees.  Dict keys are visited in SORTED order, which is what JAX does and what fixes the flat
parameter order (optimizer/vmc.py:127,265)."""
import torch


def _is_leaf(x):
    return not isinstance(x, (dict, list, tuple)) and x is not None


def leaves(t):
    if t is None:
        return []
    if _is_leaf(t):
        return [t]
    if isinstance(t, dict):
        return [l for k in sorted(t) for l in leaves(t[k])]
    return [l for v in t for l in leaves(v)]


def map(f, t, *rest):  # noqa: A001
    if t is None:
        return None
    if _is_leaf(t):
        return f(t, *rest)
    if isinstance(t, dict):
        return {k: map(f, t[k], *[r[k] for r in rest]) for k in t}
    if isinstance(t, tuple) and hasattr(t, "_fields"):
        return type(t)(*[map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t)])
    return type(t)(map(f, v, *[r[i] for r in rest]) for i, v in enumerate(t))


def ravel(t):
    """Concatenated row-major leaves = jax.flatten_util.ravel_pytree(t)[0]."""
    return torch.cat([torch.as_tensor(l).reshape(-1) for l in leaves(t)])


def paths(t, prefix=""):
    """[(path, leaf)] in leaf order."""
    if _is_leaf(t):
        return [(prefix, t)]
    out = []
    if isinstance(t, dict):
        for k in sorted(t):
            out += paths(t[k], f"{prefix}/{k}" if prefix else str(k))
    else:
        for i, v in enumerate(t):
            out += paths(v, f"{prefix}/{i}" if prefix else str(i))
    return out
