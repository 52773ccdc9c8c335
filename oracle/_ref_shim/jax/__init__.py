"""ORACLE / TEST INFRASTRUCTURE ONLY -- stand-in for the `jax` import name over torch.func.

See oracle/_ref_shim/README.md.  Exists so that the reference's own source files
(/root/reference/src/nqmc/**, which `import jax`) can be EXECUTED in the build container to
produce golden fixtures; jax itself is not installable here.  Only the API surface those files
(and the [SPEC] code blocks in /root/reference/.github/phase{2,3}-issue-body.md) use is provided.
"""
from __future__ import annotations

import torch
from torch import func as _func

from . import _config
from ._config import config  # noqa: F401  (jax.config.update("jax_enable_x64", True))
from . import numpy  # noqa: F401
from . import random  # noqa: F401
from . import lax  # noqa: F401
from . import tree  # noqa: F401

__version__ = "0.0-shim-over-torch-" + torch.__version__

Array = torch.Tensor


def grad(f, argnums=0, has_aux=False):
    """jax.grad -> torch.func.grad (reverse mode; pytree arguments supported)."""
    return _func.grad(f, argnums=argnums, has_aux=has_aux)


def hessian(f, argnums=0):
    """jax.hessian = jacfwd(jacrev(f)); torch.func.hessian is the same composition."""
    return _func.hessian(f, argnums=argnums)


def jacfwd(f, argnums=0):
    return _func.jacfwd(f, argnums=argnums)


def jacrev(f, argnums=0):
    return _func.jacrev(f, argnums=argnums)


def vmap(f, in_axes=0, out_axes=0):
    return _func.vmap(f, in_dims=in_axes, out_dims=out_axes)


def jit(f, *a, **k):   # the reference never jits; harmless identity for completeness
    return f


def _astype(self, dtype):
    return self.to(_config.to_torch_dtype(dtype))


# jnp arrays have .astype (sampler/metropolis.py:140); torch tensors do not.  Test-side only.
if not hasattr(torch.Tensor, "astype"):
    torch.Tensor.astype = _astype
