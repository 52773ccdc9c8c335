"""ORACLE (test infrastructure only) -- the reference's *own* algorithm, restated with torch.func.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this file.  PARITY UNPINNED (see oracle/layout.py).

This is restatement (i) of SURVEY section 7 step 1: a per-walker scalar log|psi| closure
driven by vmap / grad / hessian, i.e. exactly how the reference computes things --
``torch.func.{vmap,grad,hessian}`` are the one-to-one analogues of the
``jax.{vmap,grad,hessian}`` calls at

    src/nqmc/hamiltonian/molecular.py:60,63,195      (grad, dense hessian, vmap)
    src/nqmc/sampler/metropolis.py:124               (vmap(log_prob_fn))
    src/nqmc/optimizer/vmc.py:109-111,120-123        (vmap(local_energy), vmap(grad wrt params))

It is also the CPU baseline bench.py times (``cpu_baseline.kind = "port"``): the reference's
JAX path cannot run on this image (no jax/flax/optax), so this is the closest
executable statement of its cost model (dense 3N x 3N Hessian per walker, then trace).

The ansatz definitions follow .github/phase2-issue-body.md:38-66,78-113,125-152 and
.github/phase3-issue-body.md:50-83,95-122,134-169,196-219, including
backflow, which the analytic NumPy oracle does not cover.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np
import torch
from torch.func import grad, hessian, vmap

from .layout import (KIND_SIMPLE, OracleNet, OracleSystem, leaf_table, param_count)

EPSILON = 1e-10   # molecular.py:33


def _unflatten_t(sysm, net, theta: torch.Tensor) -> Dict[str, torch.Tensor]:
    out = {}
    for path, shape, off in leaf_table(sysm, net):
        n = int(np.prod(shape)) if len(shape) else 1
        out[path] = theta[off:off + n].reshape(shape)
    return out


def _safe_norm(v):
    return torch.sqrt(torch.sum(v * v, dim=-1))


def _det(M):
    """Determinant of a small (n<=5) matrix by Laplace expansion along row 0.
    torch.linalg.slogdet is NOT used: on torch 2.11 its batching rule under
    vmap(hessian(.)) returns wrong second derivatives (checked against a python loop),
    so the oracle expands the determinant explicitly; sign/log|det| are taken afterwards,
    which is what ``jnp.linalg.slogdet`` (phase2-issue-body.md:48) returns."""
    n = M.shape[0]
    if n == 1:
        return M[0, 0]
    if n == 2:
        return M[0, 0] * M[1, 1] - M[0, 1] * M[1, 0]
    out = 0.0
    cols = list(range(n))
    for j in range(n):
        minor = M[1:, cols[:j] + cols[j + 1:]]
        out = out + ((-1.0) ** j) * M[0, j] * _det(minor)
    return out


class AutodiffOracle:
    def __init__(self, sysm: OracleSystem, net: OracleNet, dtype=torch.float64):
        self.sys, self.net, self.dtype = sysm, net, dtype
        self.R = torch.tensor(sysm.R, dtype=dtype)
        self.Z = torch.tensor(sysm.Z, dtype=dtype)
        self.P = param_count(sysm, net)

    # ---------------------------------------------------------------- single-walker psi
    def _mlp(self, p, prefix, x):
        for i in range(2):
            x = torch.tanh(x @ p[f"{prefix}/Dense_{i}/kernel"] + p[f"{prefix}/Dense_{i}/bias"])
        return (x @ p[f"{prefix}/Dense_2/kernel"] + p[f"{prefix}/Dense_2/bias"]).squeeze(-1)

    def _orbital(self, p, prefix, x):
        """NeuralOrbital.__call__ (phase2-issue-body.md:83-101) for one electron position x (3,)."""
        d = _safe_norm(x[None, :] - self.R)
        f = torch.cat([x, d, torch.exp(-d)])
        return self._mlp(p, prefix, f)

    def sign_logabs(self, theta: torch.Tensor, r: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """(sign, log|psi|) for one walker r (N,3) (phase3-issue-body.md:196-219)."""
        p = _unflatten_t(self.sys, self.net, theta)
        s, net = self.sys, self.net
        N = s.n_elec
        if net.kind == KIND_SIMPLE:
            x = r.reshape(-1)                                                 # simple.py:51-68
            act = torch.tanh if net.activation == "tanh" else torch.nn.functional.silu
            for i in range(len(net.hidden)):
                x = act(x @ p[f"params/Dense_{i}/kernel"] + p[f"params/Dense_{i}/bias"])
            nh = len(net.hidden)
            mlp = (x @ p[f"params/Dense_{nh}/kernel"] + p[f"params/Dense_{nh}/bias"]).squeeze(-1)
            d = _safe_norm(r[:, None, :] - self.R[None, :, :])
            env = torch.sum(torch.log(torch.sum(torch.exp(-net.envelope_decay * d), dim=1)))  # simple.py:141-149
            return torch.ones((), dtype=r.dtype), mlp + env
        # pair geometry on *physical* coordinates
        iu = torch.triu_indices(N, N, offset=1)
        if N >= 2:
            rij = r[iu[0]] - r[iu[1]]
            dee = _safe_norm(rij)
        # backflow: x_i = r_i + sum_{j!=i} eta(r_ij)(r_i - r_j)   (phase3-issue-body.md:134-169)
        x = r
        if net.use_backflow and N >= 2:
            feat = torch.stack([dee, torch.exp(-dee)], dim=-1)
            eta = self._mlp(p, "params/backflow", feat) * torch.exp(-0.5 * dee)
            disp = torch.zeros_like(r)
            disp = disp.index_add(0, iu[0], eta[:, None] * rij)
            disp = disp.index_add(0, iu[1], -eta[:, None] * rij)
            x = r + disp
        sign = torch.ones((), dtype=r.dtype)
        logabs = torch.zeros((), dtype=r.dtype)
        for name, lo, ns in (("up", 0, s.n_up), ("down", s.n_up, s.n_dn)):
            if ns == 0:
                continue
            rows = []
            for i in range(ns):
                rows.append(torch.stack([self._orbital(p, f"params/orbitals_{name}/orbital_{k}", x[lo + i])
                                         for k in range(ns)]))
            Phi = torch.stack(rows)                                            # (n_s, n_s)
            det = _det(Phi)                                                    # phase2-issue-body.md:48
            sg, la = torch.sign(det), torch.log(torch.abs(det))
            sign = sign * sg
            logabs = logabs + la
        # Jastrow (phase2-issue-body.md:125-152)
        den = _safe_norm(r[:, None, :] - self.R[None, :, :])
        a_ee, b_ee = p["params/jastrow/a_ee"], p["params/jastrow/b_ee"]
        a_en, b_en = p["params/jastrow/a_en"], p["params/jastrow/b_en"]
        J = torch.sum(a_en * den / (1 + b_en * den))
        if N >= 2:
            J = J + torch.sum(a_ee * dee / (1 + b_ee * dee))
        logabs = logabs + J
        if net.use_cusp:
            C = -torch.sum(self.Z[None, :] * den)                              # phase3-issue-body.md:66
            for a in range(s.n_atoms):
                da = den[:, a]
                feat = torch.stack([da, da * da, torch.exp(-da)], dim=-1)
                g = torch.tanh(feat @ p[f"params/en_cusp/Dense_{2*a}/kernel"] + p[f"params/en_cusp/Dense_{2*a}/bias"])
                g = (g @ p[f"params/en_cusp/Dense_{2*a+1}/kernel"] + p[f"params/en_cusp/Dense_{2*a+1}/bias"]).squeeze(-1)
                C = C + torch.sum(da * g)                                      # :81
            if N >= 2:
                same = ((iu[0] < s.n_up) == (iu[1] < s.n_up))
                c = torch.where(same, torch.tensor(net.cusp_same, dtype=r.dtype),
                                torch.tensor(net.cusp_diff, dtype=r.dtype))
                b = p["params/ee_cusp/b_ee"]
                C = C + torch.sum(c * dee / (1 + torch.abs(b) * dee))          # :120
            logabs = logabs + C
        return sign, logabs

    def log_psi_single(self, theta, r):
        return self.sign_logabs(theta, r)[1]

    # ---------------------------------------------------------------- hamiltonian
    def kinetic_single(self, theta, r):
        """molecular.py:52-69: grad + dense Hessian + trace."""
        flat = r.reshape(-1)
        f = lambda rf: self.log_psi_single(theta, rf.reshape(r.shape))
        g = grad(f)(flat)
        H = hessian(f)(flat)
        return -0.5 * (torch.trace(H) + torch.sum(g * g))

    def potential_single(self, r):
        den = torch.clamp(_safe_norm(r[:, None, :] - self.R[None, :, :]), min=EPSILON)   # molecular.py:93-100
        v = -torch.sum(self.Z[None, :] / den)
        N = r.shape[0]
        if N >= 2:                                                                        # molecular.py:116-133
            iu = torch.triu_indices(N, N, offset=1)
            v = v + torch.sum(1.0 / _safe_norm(r[iu[0]] - r[iu[1]]))
        return v + self.sys.v_nn                                                          # molecular.py:173

    def local_energy_single(self, theta, r):
        return self.kinetic_single(theta, r) + self.potential_single(r)

    # ---------------------------------------------------------------- batched entry points
    def _t(self, a):
        return torch.as_tensor(np.asarray(a), dtype=self.dtype)

    def log_psi(self, theta, r):
        s, la = vmap(lambda x: self.sign_logabs(self._t(theta), x))(self._t(r))
        return s.numpy(), la.numpy()

    def local_energy(self, theta, r, chunk: int = 2048):
        th, rr = self._t(theta), self._t(r)
        f = vmap(lambda x: self.local_energy_single(th, x))                    # molecular.py:195
        out = [f(rr[i:i + chunk]) for i in range(0, rr.shape[0], chunk)]
        return torch.cat(out).numpy()

    def grad_log_psi(self, theta, r, chunk: int = 2048):
        th, rr = self._t(theta), self._t(r)
        f = vmap(lambda x: grad(lambda p: self.log_psi_single(p, x))(th))      # vmc.py:120-123
        out = [f(rr[i:i + chunk]) for i in range(0, rr.shape[0], chunk)]
        return torch.cat(out).numpy()

    def mh_step(self, theta, r, logp, normals, uniforms, sigma):
        """metropolis.py:115-147 with supplied streams."""
        th, rr = self._t(theta), self._t(r)
        prop = rr + sigma * self._t(normals)
        logp_prop = 2.0 * vmap(lambda x: self.log_psi_single(th, x))(prop)     # base.py:76
        ratio = logp_prop - self._t(logp)
        accept = torch.log(self._t(uniforms)) < ratio
        r_new = torch.where(accept[:, None, None], prop, rr)
        logp_new = torch.where(accept, logp_prop, self._t(logp))
        return r_new.numpy(), logp_new.numpy(), accept.numpy(), ratio.numpy()

    def walker_step(self, theta, r, logp, normals, uniforms, sigma, chunk: int = 2048):
        """One walker-step the way the reference does it: the timed unit of the CPU baseline."""
        r2, logp2, acc, _ = self.mh_step(theta, r, logp, normals, uniforms, sigma)
        e = self.local_energy(theta, r2, chunk)
        O = self.grad_log_psi(theta, r2, chunk)
        mean = e.mean()
        g = 2.0 * np.mean((e - mean)[:, None] * O, axis=0)                     # vmc.py:126-133
        return r2, logp2, acc, e, g
