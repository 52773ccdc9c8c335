"""ORACLE / TEST INFRASTRUCTURE ONLY -- `jax.lax.scan` as the Python loop it denotes."""
import torch


def _stack(ys):
    if not ys or ys[0] is None:
        return None
    y0 = ys[0]
    if isinstance(y0, torch.Tensor):
        return torch.stack(ys)
    if isinstance(y0, tuple) and hasattr(y0, "_fields"):
        return type(y0)(*[_stack([y[i] for y in ys]) for i in range(len(y0))])
    if isinstance(y0, (tuple, list)):
        return type(y0)(_stack([y[i] for y in ys]) for i in range(len(y0)))
    if isinstance(y0, dict):
        return {k: _stack([y[k] for y in ys]) for k in y0}
    return torch.stack([torch.as_tensor(y) for y in ys])


def scan(f, init, xs, length=None):
    """carry, ys = scan(f, init, xs): f(carry, x) -> (carry, y); xs sliced on its leading axis
    (sampler/metropolis.py:191,210 scan over stacked PRNG keys)."""
    carry, ys = init, []
    n = length if xs is None else len(xs)
    for i in range(n):
        carry, y = f(carry, None if xs is None else xs[i])
        ys.append(y)
    return carry, _stack(ys)
