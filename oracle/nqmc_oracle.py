"""ORACLE (test infrastructure only) -- analytic NumPy restatement of the nqmc walker loop.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this file; the product never routes through it.

PARITY UNPINNED for the antisymmetric ansatz (see oracle/layout.py header): the
reference holds no golden vectors for this path and cannot run here (no JAX).
This file is restatement (ii) of SURVEY section 7 step 1: the *analytic*
forward-Laplacian algorithm.  Restatement (i) -- the reference's own nested
autodiff algorithm -- is oracle/nqmc_autodiff.py; tests/test_oracle_consistency.py
requires (i) == (ii) to 1e-9 in float64.

Everything is batched over a leading walker axis and runs in the dtype of its
inputs (float64 by default; float32 to mimic the reference's arithmetic).

Reference lines followed
------------------------
log|psi| (kind 0)     src/nqmc/wavefunction/simple.py:41-70,124-164
log|psi| (kind 1)     .github/phase2-issue-body.md:38-66,78-113,125-152 ;
                      .github/phase3-issue-body.md:50-83,95-122,196-219 ;
                      .github/phase4-issue-body.md:92-111
kinetic energy        src/nqmc/hamiltonian/molecular.py:52-69   T = -1/2 (tr H + |g|^2)
V_en / V_ee / V_nn    src/nqmc/hamiltonian/molecular.py:87-102,116-135,173
MH step               src/nqmc/sampler/metropolis.py:115-147
gradient estimator    src/nqmc/optimizer/vmc.py:104-135
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import numpy as np

from .layout import (KIND_ANTISYMMETRIC, KIND_SIMPLE, OracleNet, OracleSystem,
                     leaf_table, param_count, unflatten)

EPSILON = 1e-10   # molecular.py:33


# --------------------------------------------------------------------------- helpers
def _tanh_rules(u, Ju, Lu):
    """value / Jacobian / Laplacian through tanh.
    u (...,H), Ju (...,D,H), Lu (...,H)."""
    h = np.tanh(u)
    tp = 1.0 - h * h
    tpp = -2.0 * h * tp
    Jh = tp[..., None, :] * Ju
    Lh = tp * Lu + tpp * np.sum(Ju * Ju, axis=-2)
    return h, Jh, Lh


def _mlp3(p: Dict[str, np.ndarray], prefix: str, f, Jf=None, Lf=None):
    """Dense-tanh-Dense-tanh-Dense(1) (phase2-issue-body.md:96-101; simple.py:63-68).
    Returns (phi, Jphi, Lphi, cache) with cache holding activations for the reverse pass."""
    W1, b1 = p[prefix + "/Dense_0/kernel"], p[prefix + "/Dense_0/bias"]
    W2, b2 = p[prefix + "/Dense_1/kernel"], p[prefix + "/Dense_1/bias"]
    W3, b3 = p[prefix + "/Dense_2/kernel"], p[prefix + "/Dense_2/bias"]
    lap = Jf is not None
    u1 = f @ W1 + b1
    if lap:
        h1, Jh1, Lh1 = _tanh_rules(u1, Jf @ W1, Lf @ W1)
    else:
        h1 = np.tanh(u1)
    u2 = h1 @ W2 + b2
    if lap:
        h2, Jh2, Lh2 = _tanh_rules(u2, Jh1 @ W2, Lh1 @ W2)
    else:
        h2 = np.tanh(u2)
    phi = (h2 @ W3)[..., 0] + b3[0]
    cache = (f, h1, h2)
    if lap:
        return phi, (Jh2 @ W3)[..., 0], (Lh2 @ W3)[..., 0], cache
    return phi, None, None, cache


def _mlp3_backward(p, prefix, cache, seed, grads: Dict[str, np.ndarray]):
    """Reverse pass for d(sum_rows seed*phi)/dtheta, *per walker* (leading axis kept).
    cache arrays have shape (W, rows, .); seed (W, rows).  grads[path] gets (W, *shape)."""
    f, h1, h2 = cache
    W2 = p[prefix + "/Dense_1/kernel"]
    W3 = p[prefix + "/Dense_2/kernel"]
    grads[prefix + "/Dense_2/bias"] = seed.sum(-1)[:, None]
    grads[prefix + "/Dense_2/kernel"] = np.einsum("wr,wrh->wh", seed, h2)[..., None]
    d2 = seed[..., None] * W3[:, 0] * (1.0 - h2 * h2)
    grads[prefix + "/Dense_1/bias"] = d2.sum(-2)
    grads[prefix + "/Dense_1/kernel"] = np.einsum("wrk,wrj->wkj", h1, d2)
    d1 = (d2 @ W2.T) * (1.0 - h1 * h1)
    grads[prefix + "/Dense_0/bias"] = d1.sum(-2)
    grads[prefix + "/Dense_0/kernel"] = np.einsum("wrf,wrj->wfj", f, d1)


def _act(name, u):
    """(sigma, sigma', sigma'') for simple.py:56-58's two activations."""
    if name == "tanh":
        h = np.tanh(u)
        tp = 1.0 - h * h
        return h, tp, -2.0 * h * tp
    if name == "silu":
        s = 1.0 / (1.0 + np.exp(-u))
        return u * s, s * (1.0 + u * (1.0 - s)), s * (1.0 - s) * (2.0 + u * (1.0 - 2.0 * s))
    raise ValueError(f"Unknown activation: {name}")


def _mlpN(p, prefix, n_hidden, act, x, lap: bool):
    """SimpleMLP (simple.py:51-68) with any number of hidden layers: value, Jacobian (W,D,.), Laplacian.
    cache = per-layer (input, sigma'(u)) for the reverse pass."""
    Wn, D = x.shape
    h = x
    J = np.broadcast_to(np.eye(D, dtype=x.dtype), (Wn, D, D)) if lap else None
    L = np.zeros((Wn, D), dtype=x.dtype) if lap else None
    cache = []
    for i in range(n_hidden):
        Wi, bi = p[f"{prefix}/Dense_{i}/kernel"], p[f"{prefix}/Dense_{i}/bias"]
        u = h @ Wi + bi
        a, a1, a2 = _act(act, u)
        cache.append((h, a1))
        if lap:
            Ju = J @ Wi
            L = a1 * (L @ Wi) + a2 * np.sum(Ju * Ju, axis=-2)
            J = a1[:, None, :] * Ju
        h = a
    Wo, bo = p[f"{prefix}/Dense_{n_hidden}/kernel"], p[f"{prefix}/Dense_{n_hidden}/bias"]
    cache.append((h, None))
    val = (h @ Wo)[:, 0] + bo[0]
    if lap:
        return val, (J @ Wo)[..., 0], (L @ Wo)[:, 0], cache
    return val, None, None, cache


def _mlpN_backward(p, prefix, n_hidden, cache, grads):
    """Per-walker d val / d theta for _mlpN."""
    h_last, _ = cache[n_hidden]
    Wn = h_last.shape[0]
    grads[f"{prefix}/Dense_{n_hidden}/bias"] = np.ones((Wn, 1), dtype=h_last.dtype)
    grads[f"{prefix}/Dense_{n_hidden}/kernel"] = h_last[..., None]
    back = np.broadcast_to(p[f"{prefix}/Dense_{n_hidden}/kernel"][:, 0], h_last.shape)
    for i in range(n_hidden - 1, -1, -1):
        h_in, a1 = cache[i]
        d = back * a1
        grads[f"{prefix}/Dense_{i}/bias"] = d
        grads[f"{prefix}/Dense_{i}/kernel"] = h_in[:, :, None] * d[:, None, :]
        back = d @ p[f"{prefix}/Dense_{i}/kernel"].T


def orbital_features(x, R):
    """phase2-issue-body.md:83-93: f = [x(3), |x-R_A| (A), exp(-|x-R_A|) (A)].
    x (W,n,3); returns f (W,n,F), Jf (W,n,3,F), Lf (W,n,F) (Appendix C.1 input rules)."""
    W, n, _ = x.shape
    A = R.shape[0]
    diff = x[:, :, None, :] - R[None, None, :, :].astype(x.dtype)   # (W,n,A,3)
    d = np.sqrt(np.sum(diff * diff, axis=-1))                        # (W,n,A)
    e = np.exp(-d)
    uhat = diff / d[..., None]
    F = 3 + 2 * A
    f = np.concatenate([x, d, e], axis=-1)
    Jf = np.zeros((W, n, 3, F), dtype=x.dtype)
    Jf[:, :, 0, 0] = 1.0
    Jf[:, :, 1, 1] = 1.0
    Jf[:, :, 2, 2] = 1.0
    Jf[:, :, :, 3:3 + A] = np.swapaxes(uhat, -1, -2)
    Jf[:, :, :, 3 + A:] = np.swapaxes(-e[..., None] * uhat, -1, -2)
    Lf = np.zeros((W, n, F), dtype=x.dtype)
    Lf[:, :, 3:3 + A] = 2.0 / d
    Lf[:, :, 3 + A:] = e * (1.0 - 2.0 / d)
    return f, Jf, Lf


def _pade(r, a, b):
    """u = a r/(1+b r), u', u'' (phase2-issue-body.md:141-142)."""
    q = 1.0 / (1.0 + b * r)
    return a * r * q, a * q * q, -2.0 * a * b * q * q * q


def _pair_terms(r):
    """r (W,N,3) -> i<j index arrays, r_ij (W,npair), rhat (W,npair,3)."""
    N = r.shape[1]
    ii, jj = np.triu_indices(N, k=1)
    diff = r[:, ii, :] - r[:, jj, :]
    d = np.sqrt(np.sum(diff * diff, axis=-1))
    return ii, jj, d, diff / d[..., None]


class Oracle:
    """Analytic evaluator for one (system, net).  All arrays batched over walkers."""

    def __init__(self, sysm: OracleSystem, net: OracleNet):
        self.sys, self.net = sysm, net
        self.P = param_count(sysm, net)
        self.table = leaf_table(sysm, net)

    # ------------------------------------------------------------------ kind 1 pieces
    def _spin_slices(self):
        return [("up", slice(0, self.sys.n_up), self.sys.n_up),
                ("down", slice(self.sys.n_up, self.sys.n_elec), self.sys.n_dn)]

    def _orbital_matrices(self, p, r, lap: bool):
        """Phi_s[w,i,k] (+ grad (W,i,k,3) and Laplacian) per spin (phase4-issue-body.md:97-105)."""
        out = []
        for name, sl, ns in self._spin_slices():
            if ns == 0:
                out.append(None)
                continue
            x = r[:, sl, :]
            f, Jf, Lf = orbital_features(x, self.sys.R)
            phis, gs, ls, caches = [], [], [], []
            for k in range(ns):
                phi, g, l, cache = _mlp3(p, f"params/orbitals_{name}/orbital_{k}", f,
                                         Jf if lap else None, Lf if lap else None)
                phis.append(phi); gs.append(g); ls.append(l); caches.append(cache)
            Phi = np.stack(phis, axis=-1)
            G = np.stack(gs, axis=-2) if lap else None           # (W,i,k,3)
            L = np.stack(ls, axis=-1) if lap else None
            out.append((Phi, G, L, caches))
        return out

    def _jastrow(self, p, r, lap: bool):
        """phase2-issue-body.md:125-152 on physical r; returns J, grad (W,N,3), lap (W,)."""
        a_ee, b_ee = p["params/jastrow/a_ee"], p["params/jastrow/b_ee"]
        a_en, b_en = p["params/jastrow/a_en"], p["params/jastrow/b_en"]
        Wn, N, _ = r.shape
        J = np.zeros(Wn, dtype=r.dtype)
        g = np.zeros_like(r)
        l = np.zeros(Wn, dtype=r.dtype)
        extra = {}
        if N >= 2:
            ii, jj, d, rh = _pair_terms(r)
            u, up, upp = _pade(d, a_ee, b_ee)
            J += u.sum(-1)
            if lap:
                np.add.at(g, (slice(None), ii), up[..., None] * rh)
                np.add.at(g, (slice(None), jj), -up[..., None] * rh)
                l += 2.0 * (upp + 2.0 * up / d).sum(-1)
            extra["ee"] = d
        diff = r[:, :, None, :] - self.sys.R[None, None].astype(r.dtype)
        den = np.sqrt(np.sum(diff * diff, -1))
        u, up, upp = _pade(den, a_en, b_en)
        J += u.sum((-1, -2))
        if lap:
            g += np.sum(up[..., None] * diff / den[..., None], axis=2)
            l += (upp + 2.0 * up / den).sum((-1, -2))
        extra["en"] = den
        return J, g, l, extra

    def _cusps(self, p, r, lap: bool):
        """phase3-issue-body.md:50-83 (e-n) and :95-122 (e-e), on physical r."""
        Wn, N, _ = r.shape
        A = self.sys.n_atoms
        C = np.zeros(Wn, dtype=r.dtype)
        g = np.zeros_like(r)
        l = np.zeros(Wn, dtype=r.dtype)
        diff = r[:, :, None, :] - self.sys.R[None, None].astype(r.dtype)
        d = np.sqrt(np.sum(diff * diff, -1))                 # (W,N,A)
        uh = diff / d[..., None]
        Z = self.sys.Z.astype(r.dtype)
        # exact cusp  -sum Z d
        C += -(Z * d).sum((-1, -2))
        if lap:
            g += -(Z[None, None, :, None] * uh).sum(2)
            l += -(Z * 2.0 / d).sum((-1, -2))
        # smooth correction  sum_A sum_i d * g_A(d),  g_A = Dense1(tanh(Dense16([d,d^2,e^-d])))
        caches = []
        for a in range(A):
            W0 = p[f"params/en_cusp/Dense_{2*a}/kernel"]; b0 = p[f"params/en_cusp/Dense_{2*a}/bias"]
            W1 = p[f"params/en_cusp/Dense_{2*a+1}/kernel"]; b1 = p[f"params/en_cusp/Dense_{2*a+1}/bias"]
            da = d[:, :, a]
            ea = np.exp(-da)
            feat = np.stack([da, da * da, ea], -1)            # (W,N,3)
            dfeat = np.stack([np.ones_like(da), 2 * da, -ea], -1)     # d/dd
            ddfeat = np.stack([np.zeros_like(da), 2 * np.ones_like(da), ea], -1)
            z = feat @ W0 + b0
            t = np.tanh(z)
            gval = (t @ W1)[..., 0] + b1[0]
            C += (da * gval).sum(-1)
            caches.append((feat, t, da))
            if lap:
                tp = 1 - t * t
                zp = dfeat @ W0
                zpp = ddfeat @ W0
                gp = ((tp * zp) @ W1)[..., 0]
                gpp = ((tp * zpp - 2 * t * tp * zp * zp) @ W1)[..., 0]
                # q(d) = d g(d): q' = g + d g', q'' = 2 g' + d g''
                qp = gval + da * gp
                qpp = 2 * gp + da * gpp
                g += qp[..., None] * uh[:, :, a, :]
                l += (qpp + 2.0 * qp / da).sum(-1)
        # e-e cusp
        extra = {"en_caches": caches}
        if N >= 2:
            b = p["params/ee_cusp/b_ee"]
            ii, jj, dee, rh = _pair_terms(r)
            same = ((ii < self.sys.n_up) == (jj < self.sys.n_up))
            c = np.where(same, self.net.cusp_same, self.net.cusp_diff).astype(r.dtype)
            u, up, upp = _pade(dee, c, np.abs(b))
            C += u.sum(-1)
            if lap:
                np.add.at(g, (slice(None), ii), up[..., None] * rh)
                np.add.at(g, (slice(None), jj), -up[..., None] * rh)
                l += 2.0 * (upp + 2.0 * up / dee).sum(-1)
            extra["ee"] = (dee, c)
        return C, g, l, extra

    # ------------------------------------------------------------------ public API
    def log_psi(self, theta, r) -> Tuple[np.ndarray, np.ndarray]:
        """(sign, log|psi|) for r (W,N,3)."""
        s, la, _ = self._forward(theta, r, lap=False)
        return s, la

    def _forward(self, theta, r, lap: bool):
        p = unflatten(self.sys, self.net, np.asarray(theta, dtype=r.dtype))
        if self.net.kind == KIND_SIMPLE:
            return self._forward_simple(p, r, lap)
        if self.net.use_backflow:
            raise NotImplementedError("analytic oracle: backflow is covered by oracle/nqmc_autodiff.py")
        Wn, N, _ = r.shape
        mats = self._orbital_matrices(p, r, lap)
        sign = np.ones(Wn, dtype=r.dtype)
        logabs = np.zeros(Wn, dtype=r.dtype)
        g = np.zeros_like(r)
        lapsum = np.zeros(Wn, dtype=r.dtype)          # sum_i (Lap_i log|det| + |grad_i log|det||^2)
        invs = []
        for (name, sl, ns), m in zip(self._spin_slices(), mats):
            if m is None:
                invs.append(None)
                continue
            Phi, G, L, caches = m
            sg, la = np.linalg.slogdet(Phi)             # phase2-issue-body.md:48
            sign = sign * sg
            logabs = logabs + la
            inv = np.linalg.inv(Phi)                    # inv[w,k,i]
            invs.append(inv)
            if lap:
                invT = np.swapaxes(inv, -1, -2)         # [w,i,k]
                g[:, sl, :] += np.einsum("wik,wikd->wid", invT, G)     # Appendix C.2
                lapsum += np.einsum("wik,wik->w", invT, L)
        J, gJ, lJ, _ = self._jastrow(p, r, lap)
        logabs = logabs + J
        gdet = g.copy()
        if lap:
            g = g + gJ
        lother = lJ
        if self.net.use_cusp:
            C, gC, lC, _ = self._cusps(p, r, lap)
            logabs = logabs + C
            if lap:
                g = g + gC
                lother = lother + lC
        aux = None
        if lap:
            # Lap log|det| = lapsum - |grad log|det||^2 ; total T = -1/2 (Lap logPsi + |grad logPsi|^2)
            lap_log = lapsum - np.sum(gdet * gdet, axis=(-1, -2)) + lother
            aux = (g, lap_log, mats, invs, p)
        else:
            aux = (None, None, mats, invs, p)
        return sign, logabs, aux

    def _forward_simple(self, p, r, lap: bool):
        """simple.py:124-164: MLP(flatten r) + sum_i log sum_A exp(-alpha |r_i-R_A|)."""
        Wn, N, _ = r.shape
        D = 3 * N
        x = r.reshape(Wn, D)
        mlp, Jm, Lm, cache = _mlpN(p, "params", len(self.net.hidden), self.net.activation, x, lap)
        alpha = self.net.envelope_decay
        diff = r[:, :, None, :] - self.sys.R[None, None].astype(r.dtype)
        d = np.sqrt(np.sum(diff * diff, -1))
        ex = np.exp(-alpha * d)
        S = ex.sum(-1)                                   # (W,N)
        env = np.log(S).sum(-1)
        logabs = mlp + env
        sign = np.ones(Wn, dtype=r.dtype)
        aux = (None, None, cache, None, p)
        if lap:
            uh = diff / d[..., None]
            gS = (-alpha * ex[..., None] * uh).sum(2)                      # (W,N,3)
            lS = (ex * (alpha * alpha - 2.0 * alpha / d)).sum(-1)          # (W,N)
            genv = gS / S[..., None]
            lenv = (lS / S - np.sum(genv * genv, -1)).sum(-1)
            g = Jm.reshape(Wn, N, 3) + genv
            aux = (g, Lm + lenv, cache, None, p)
        return sign, logabs, aux

    # potentials -----------------------------------------------------------------
    def potential_en(self, r):
        """molecular.py:87-102."""
        diff = r[:, :, None, :] - self.sys.R[None, None].astype(r.dtype)
        d = np.maximum(np.sqrt(np.sum(diff * diff, -1)), r.dtype.type(EPSILON))
        return -(self.sys.Z.astype(r.dtype) / d).sum((-1, -2))

    def potential_ee(self, r):
        """molecular.py:116-135 (0 when fewer than two electrons)."""
        if r.shape[1] < 2:
            return np.zeros(r.shape[0], dtype=r.dtype)
        _, _, d, _ = _pair_terms(r)
        return (1.0 / d).sum(-1)

    def local_energy(self, theta, r):
        """molecular.py:159-175: E_L = T + V_en + V_ee + V_nn.  Returns (E_L, parts dict, sign, logabs)."""
        sign, logabs, aux = self._forward(theta, r, lap=True)
        g, lap_log = aux[0], aux[1]
        T = -0.5 * (lap_log + np.sum(g * g, axis=(-1, -2)))        # molecular.py:67
        ven = self.potential_en(r)
        vee = self.potential_ee(r)
        e = T + ven + vee + r.dtype.type(self.sys.v_nn)
        return e, {"T": T, "V_en": ven, "V_ee": vee, "V_nn": self.sys.v_nn}, sign, logabs

    # parameter gradient ---------------------------------------------------------
    def grad_log_psi(self, theta, r):
        """O[w,p] = d log|psi(r_w)| / d theta_p (vmc.py:120-123), analytic reverse pass."""
        sign, logabs, aux = self._forward(theta, r, lap=False)
        p = aux[4]
        Wn, N, _ = r.shape
        grads: Dict[str, np.ndarray] = {}
        if self.net.kind == KIND_SIMPLE:
            _mlpN_backward(p, "params", len(self.net.hidden), aux[2], grads)
        else:
            mats, invs = aux[2], aux[3]
            for (name, sl, ns), m, inv in zip(self._spin_slices(), mats, invs):
                if m is None:
                    continue
                caches = m[3]
                for k in range(ns):
                    # d log|det Phi| / d Phi[i,k] = inv[k,i]   (Appendix C.4)
                    _mlp3_backward(p, f"params/orbitals_{name}/orbital_{k}", caches[k],
                                   inv[:, k, :], grads)
            # Jastrow scalars
            a_ee, b_ee = p["params/jastrow/a_ee"], p["params/jastrow/b_ee"]
            a_en, b_en = p["params/jastrow/a_en"], p["params/jastrow/b_en"]
            diff = r[:, :, None, :] - self.sys.R[None, None].astype(r.dtype)
            den = np.sqrt(np.sum(diff * diff, -1))
            q = 1.0 / (1.0 + b_en * den)
            grads["params/jastrow/a_en"] = (den * q).sum((-1, -2))
            grads["params/jastrow/b_en"] = (-a_en * den * den * q * q).sum((-1, -2))
            if N >= 2:
                _, _, dee, _ = _pair_terms(r)
                q = 1.0 / (1.0 + b_ee * dee)
                grads["params/jastrow/a_ee"] = (dee * q).sum(-1)
                grads["params/jastrow/b_ee"] = (-a_ee * dee * dee * q * q).sum(-1)
            else:
                grads["params/jastrow/a_ee"] = np.zeros(Wn, dtype=r.dtype)
                grads["params/jastrow/b_ee"] = np.zeros(Wn, dtype=r.dtype)
            if self.net.use_cusp:
                for a in range(self.sys.n_atoms):
                    W1 = p[f"params/en_cusp/Dense_{2*a+1}/kernel"]
                    da = den[:, :, a]
                    feat = np.stack([da, da * da, np.exp(-da)], -1)
                    z = feat @ p[f"params/en_cusp/Dense_{2*a}/kernel"] + p[f"params/en_cusp/Dense_{2*a}/bias"]
                    t = np.tanh(z)
                    grads[f"params/en_cusp/Dense_{2*a+1}/bias"] = da.sum(-1)[:, None]
                    grads[f"params/en_cusp/Dense_{2*a+1}/kernel"] = np.einsum("wn,wnh->wh", da, t)[..., None]
                    dz = da[..., None] * W1[:, 0] * (1 - t * t)
                    grads[f"params/en_cusp/Dense_{2*a}/bias"] = dz.sum(-2)
                    grads[f"params/en_cusp/Dense_{2*a}/kernel"] = np.einsum("wnf,wnh->wfh", feat, dz)
                if N >= 2:
                    b = p["params/ee_cusp/b_ee"]
                    ii, jj, dee, _ = _pair_terms(r)
                    same = ((ii < self.sys.n_up) == (jj < self.sys.n_up))
                    c = np.where(same, self.net.cusp_same, self.net.cusp_diff).astype(r.dtype)
                    q = 1.0 / (1.0 + np.abs(b) * dee)
                    grads["params/ee_cusp/b_ee"] = np.sign(b) * (-c * dee * dee * q * q).sum(-1)
                else:
                    grads["params/ee_cusp/b_ee"] = np.zeros(Wn, dtype=r.dtype)
        O = np.zeros((Wn, self.P), dtype=r.dtype)
        for path, shape, off in self.table:
            n = int(np.prod(shape)) if len(shape) else 1
            O[:, off:off + n] = grads[path].reshape(Wn, n)
        return O

    # sampler ---------------------------------------------------------------------
    def mh_step(self, theta, r, logp, normals, uniforms, sigma):
        """metropolis.py:115-147 with externally supplied streams.
        log_prob = 2 log|psi| (base.py:76).  Returns r', logp', accept(bool), log_ratio."""
        dt = r.dtype.type
        prop = r + dt(sigma) * normals.astype(r.dtype)
        _, la = self.log_psi(theta, prop)
        logp_prop = dt(2.0) * la
        ratio = logp_prop - logp
        with np.errstate(divide="ignore"):
            logu = np.log(uniforms.astype(r.dtype))
        accept = logu < ratio                                   # strict (metropolis.py:133)
        r_new = np.where(accept[:, None, None], prop, r)
        logp_new = np.where(accept, logp_prop, logp)
        return r_new, logp_new, accept, ratio

    # estimator -------------------------------------------------------------------
    def energy_and_gradient(self, theta, r):
        """vmc.py:104-135 on a flat batch r (n_total,N,3):
        returns (grad (P,), mean_energy, energy_std=std(ddof=0)/sqrt(n), E_L, O)."""
        e, _, _, _ = self.local_energy(theta, r)
        n = e.shape[0]
        mean = e.mean()
        std = e.std() / np.sqrt(n)
        O = self.grad_log_psi(theta, r)
        grad = 2.0 * np.mean((e - mean)[:, None] * O, axis=0)
        return grad, mean, std, e, O

    def walker_step(self, theta, r, logp, normals, uniforms, sigma):
        """One walker-step of SURVEY section 8(d): MH move, then E_L and O at the new position.
        Returns dict with the raw sums the device path accumulates."""
        r2, logp2, acc, ratio = self.mh_step(theta, r, logp, normals, uniforms, sigma)
        e, parts, sign, la = self.local_energy(theta, r2)
        O = self.grad_log_psi(theta, r2)
        return {"r": r2, "logp": logp2, "accept": acc, "ratio": ratio, "e_l": e,
                "sum_e": e.sum(), "sum_e2": (e * e).sum(), "sum_o": O.sum(0),
                "sum_eo": (e[:, None] * O).sum(0), "O": O, "sign": sign, "logabs": la}
