"""dtype policy of the stand-in: float32 like stock JAX, float64 after
jax.config.update("jax_enable_x64", True) (JAX's own switch)."""
import torch


class _Config:
    jax_enable_x64 = False

    def update(self, name, value):
        if name != "jax_enable_x64":
            raise KeyError(name)
        self.jax_enable_x64 = bool(value)


config = _Config()


def float_dtype():
    return torch.float64 if config.jax_enable_x64 else torch.float32


class _DT:
    """jnp.float32 / jnp.float64 / jnp.int32 tokens; callable like the JAX scalar types."""

    def __init__(self, t):
        self.t = t

    def __call__(self, x):
        return torch.as_tensor(x, dtype=self.t)


def to_torch_dtype(d):
    if isinstance(d, _DT):
        return d.t
    if isinstance(d, torch.dtype):
        return d
    if d is float:
        return float_dtype()
    if d is int:
        return torch.int64
    if d is bool:
        return torch.bool
    import numpy as np
    return {np.dtype("float32"): torch.float32, np.dtype("float64"): torch.float64,
            np.dtype("int32"): torch.int32, np.dtype("int64"): torch.int64, np.dtype("bool"): torch.bool}[np.dtype(d)]
