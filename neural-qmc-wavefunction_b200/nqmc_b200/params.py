"""Parameter pytrees <-> the flat theta vector of the C ABI.

Params are nested dicts of float32 NumPy arrays shaped like the reference's flax trees
(src/nqmc/wavefunction/simple.py:166-176; [SPEC] .github/phase2-issue-body.md:105-111,137-147).
The flat order is sorted-key, row-major leaves -- what ``jax.tree.leaves`` yields
(src/nqmc/optimizer/vmc.py:265) -- and is the order libnqmc_b200 reports via nqmc_leaf_info.
"""
from __future__ import annotations

from typing import Any, Dict, Iterator, List, Tuple

import numpy as np

Params = Dict[str, Any]


def tree_leaves_with_path(tree: Params, prefix: str = "") -> Iterator[Tuple[str, np.ndarray]]:
    for key in sorted(tree.keys()):
        v = tree[key]
        path = f"{prefix}/{key}" if prefix else key
        if isinstance(v, dict):
            yield from tree_leaves_with_path(v, path)
        else:
            yield path, np.asarray(v)


def tree_leaves(tree: Params) -> List[np.ndarray]:
    return [v for _, v in tree_leaves_with_path(tree)]


def tree_map(fn, tree: Params, *rest: Params) -> Params:
    out = {}
    for key in tree.keys():
        if isinstance(tree[key], dict):
            out[key] = tree_map(fn, tree[key], *[r[key] for r in rest])
        else:
            out[key] = fn(tree[key], *[r[key] for r in rest])
    return out


def ravel(tree: Params, dtype=np.float32) -> np.ndarray:
    leaves = tree_leaves(tree)
    if not leaves:
        return np.zeros(0, dtype=dtype)
    return np.concatenate([np.asarray(l, dtype=dtype).reshape(-1) for l in leaves])


def unravel(theta: np.ndarray, like: Params, dtype=None) -> Params:
    """Inverse of ravel for a tree with the structure/shapes of ``like``."""
    pos = [0]

    def build(t):
        out = {}
        for key in sorted(t.keys()):
            v = t[key]
            if isinstance(v, dict):
                out[key] = build(v)
            else:
                v = np.asarray(v)
                n = v.size
                out[key] = np.asarray(theta[pos[0]:pos[0] + n], dtype=dtype or v.dtype).reshape(v.shape)
                pos[0] += n
        return out

    res = build(like)
    if pos[0] != np.asarray(theta).size:
        raise ValueError("theta length does not match the parameter tree")
    # keep the caller's key order
    def reorder(src, ref):
        return {k: (reorder(src[k], ref[k]) if isinstance(ref[k], dict) else src[k]) for k in ref.keys()}
    return reorder(res, like)
