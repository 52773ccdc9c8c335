"""Golden vectors for Stochastic Reconfiguration, produced by EXECUTING the reference's own definition.

The reference ships no ``src/nqmc/optimizer/sr.py``; its definition of SR is the [SPEC] code block
``/root/reference/.github/phase6-issue-body.md:140-181``.  This script extracts that block verbatim, executes it under
the jax stand-in (``oracle/_ref_shim``, jax.numpy over torch, float64) on seeded inputs and writes
``tests/golden/ref_sr_golden.npz``.  Run here (the GPU box has no /root/reference):
    python tests/golden/make_ref_sr_golden.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [os.path.join(ROOT, "oracle", "_ref_shim")]

import numpy as np  # noqa: E402
import torch  # noqa: E402
import jax  # noqa: E402  (the stand-in)
import jax.numpy as jnp  # noqa: E402

BLOCK = (".github/phase6-issue-body.md", 140, 181)


def load():
    rel, lo, hi = BLOCK
    lines = open(os.path.join(REF, rel), encoding="utf-8").read().split("\n")
    assert lines[lo - 2].startswith("```python") and lines[hi].startswith("```"), (lines[lo - 2], lines[hi])
    ns = {"jax": jax, "jnp": jnp}
    exec(compile("\n".join(lines[lo - 1:hi]), f"{rel}:{lo}-{hi}", "exec"), ns)
    return ns["StochasticReconfiguration"]


def problem(P, n, seed):
    rng = np.random.default_rng(seed)
    mix = np.eye(P) + 0.3 * rng.standard_normal((P, P)) / np.sqrt(P)
    g = (rng.standard_normal((n, P)) @ mix + 2.0 * rng.standard_normal(P)).astype(np.float32)
    e = (-1.1 + 0.3 * rng.standard_normal(n) + 0.05 * g[:, 0]).astype(np.float32)
    return g, e


def main():
    jax.config.update("jax_enable_x64", True)     # float64 run of the reference code: the golden is the exact answer
    SR = load()
    out = {}
    for name, P, n, lr, reg in [("a", 24, 256, 1e-2, 1e-4), ("b", 64, 500, 5e-2, 1e-3), ("c", 128, 640, 1e-2, 1e-2)]:
        g, e = problem(P, n, P + n)
        upd = SR(learning_rate=lr, regularization=reg).compute_update(jnp.array(g.astype(np.float64)), jnp.array(e.astype(np.float64)))
        out[f"{name}_grads"], out[f"{name}_energies"] = g, e
        out[f"{name}_update"] = np.asarray(upd.detach().cpu().numpy(), dtype=np.float64)
        out[f"{name}_lr_reg"] = np.array([lr, reg])
    np.savez_compressed(os.path.join(HERE, "ref_sr_golden.npz"), **out)
    print("wrote ref_sr_golden.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
