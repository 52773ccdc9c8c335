"""ORACLE (test infrastructure only) -- analytic NumPy restatement of the *backflow* ansatz.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this file.  PARITY UNPINNED (see oracle/layout.py).

Extends oracle/nqmc_oracle.py with the BackflowTransform of .github/phase3-issue-body.md:134-169,
   x_i = r_i + sum_{j != i} eta(r_ij) (r_i - r_j),   eta(r) = MLP([r, e^-r]; 16,16,1; tanh) * e^{-r/2},
orbitals evaluated at x, Jastrow / cusps on the physical r (phase3-issue-body.md:196-219), using the
orbital-Hessian chain rule of SURVEY Appendix C.3:
   d_mu log|det|  = tr(Phi^-1 d_mu Phi)
   d2_mu log|det| = tr(Phi^-1 d2_mu Phi) - tr((Phi^-1 d_mu Phi)^2)
   d_mu Phi[i,k]  = g_ik . d_mu x_i
   sum_mu d2_mu Phi[i,k] = sum_ab Hess_ik[a,b] (sum_mu d_mu x_ia d_mu x_ib) + g_ik . Lap x_i
tests/test_oracle.py requires this to equal the torch.func autodiff restatement to 1e-9 in float64.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from .layout import unflatten
from .nqmc_oracle import Oracle, _mlp3, _mlp3_backward, _pair_terms


def _tanh_hess_rules(u, Ju, Hu):
    """value / gradient (..,3,H) / Hessian (..,3,3,H) through tanh."""
    h = np.tanh(u)
    tp = 1.0 - h * h
    tpp = -2.0 * h * tp
    Jh = tp[..., None, :] * Ju
    Hh = tp[..., None, None, :] * Hu + tpp[..., None, None, :] * Ju[..., :, None, :] * Ju[..., None, :, :]
    return h, Jh, Hh


def orbital_features_hess(x, R):
    """f (W,n,F), Jf (W,n,3,F), Hf (W,n,3,3,F) of [x, |x-R_A|, exp(-|x-R_A|)] w.r.t. x."""
    Wn, n, _ = x.shape
    A = R.shape[0]
    diff = x[:, :, None, :] - R[None, None].astype(x.dtype)
    d = np.sqrt(np.sum(diff * diff, -1))
    e = np.exp(-d)
    u = diff / d[..., None]                                   # (W,n,A,3)
    F = 3 + 2 * A
    f = np.concatenate([x, d, e], -1)
    Jf = np.zeros((Wn, n, 3, F), dtype=x.dtype)
    Hf = np.zeros((Wn, n, 3, 3, F), dtype=x.dtype)
    for a in range(3):
        Jf[:, :, a, a] = 1.0
    uu = u[..., :, None] * u[..., None, :]                    # (W,n,A,3,3)
    I = np.eye(3, dtype=x.dtype)
    Hd = (I - uu) / d[..., None, None]
    He = e[..., None, None] * (uu * (1.0 + 1.0 / d)[..., None, None] - I / d[..., None, None])
    Jf[:, :, :, 3:3 + A] = np.swapaxes(u, -1, -2)
    Jf[:, :, :, 3 + A:] = np.swapaxes(-e[..., None] * u, -1, -2)
    Hf[..., 3:3 + A] = np.moveaxis(Hd, 2, -1)
    Hf[..., 3 + A:] = np.moveaxis(He, 2, -1)
    return f, Jf, Hf


def _mlp3_hess(p, prefix, f, Jf, Hf):
    W1, b1 = p[prefix + "/Dense_0/kernel"], p[prefix + "/Dense_0/bias"]
    W2, b2 = p[prefix + "/Dense_1/kernel"], p[prefix + "/Dense_1/bias"]
    W3, b3 = p[prefix + "/Dense_2/kernel"], p[prefix + "/Dense_2/bias"]
    h1, J1, H1 = _tanh_hess_rules(f @ W1 + b1, Jf @ W1, Hf @ W1)
    h2, J2, H2 = _tanh_hess_rules(h1 @ W2 + b2, J1 @ W2, H1 @ W2)
    return (h2 @ W3)[..., 0] + b3[0], (J2 @ W3)[..., 0], (H2 @ W3)[..., 0], (f, h1, h2)


class BackflowOracle(Oracle):
    """Analytic evaluator for nets with use_backflow=True (cusps optional)."""

    def _backflow(self, p, r, derivs: bool):
        Wn, N, _ = r.shape
        ii, jj, d, uh = _pair_terms(r)                         # d (W,P), uh (W,P,3) = (r_i - r_j)/d
        feat = np.stack([d, np.exp(-d)], -1)
        Jf = np.stack([np.ones_like(d), -np.exp(-d)], -1)[..., None, :]
        Lf = np.stack([np.zeros_like(d), np.exp(-d)], -1)
        m, mp, mpp, cache = _mlp3(p, "params/backflow", feat, Jf, Lf)
        mp = mp[..., 0]
        E = np.exp(-0.5 * d)
        eta = m * E
        x = r.copy()
        disp = eta[..., None] * (uh * d[..., None])
        np.add.at(x, (slice(None), ii), disp)
        np.add.at(x, (slice(None), jj), -disp)
        out = {"x": x, "ii": ii, "jj": jj, "d": d, "uh": uh, "cache": cache, "E": E}
        if derivs:
            etap = (mp - 0.5 * m) * E
            etapp = (mpp - mp + 0.25 * m) * E
            kappa = etap * d
            Jx = np.zeros((Wn, N, 3, N, 3), dtype=r.dtype)
            I = np.eye(3, dtype=r.dtype)
            for e in range(N):
                Jx[:, e, :, e, :] = I
            uu = uh[..., :, None] * uh[..., None, :]
            blk = eta[..., None, None] * I + kappa[..., None, None] * uu        # d x_i / d r_i contribution of pair
            for q, (i, j) in enumerate(zip(ii, jj)):
                Jx[:, i, :, i, :] += blk[:, q]
                Jx[:, i, :, j, :] -= blk[:, q]
                Jx[:, j, :, j, :] += blk[:, q]
                Jx[:, j, :, i, :] -= blk[:, q]
            lapx = np.zeros((Wn, N, 3), dtype=r.dtype)
            lf = 2.0 * (d * etapp + 4.0 * etap)[..., None] * uh
            np.add.at(lapx, (slice(None), ii), lf)
            np.add.at(lapx, (slice(None), jj), -lf)
            out.update(Jx=Jx, lapx=lapx)
        return out

    def _forward(self, theta, r, lap: bool):
        p = unflatten(self.sys, self.net, np.asarray(theta, dtype=r.dtype))
        Wn, N, _ = r.shape
        bf = self._backflow(p, r, lap)
        x = bf["x"]
        sign = np.ones(Wn, dtype=r.dtype)
        logabs = np.zeros(Wn, dtype=r.dtype)
        gdet = np.zeros((Wn, N, 3), dtype=r.dtype)
        lap_det = np.zeros(Wn, dtype=r.dtype)
        v_all = np.zeros((Wn, N, 3), dtype=r.dtype)
        mats, invs = [], []
        for name, sl, ns in self._spin_slices():
            if ns == 0:
                mats.append(None); invs.append(None)
                continue
            xs = x[:, sl, :]
            if lap:
                f, Jf, Hf = orbital_features_hess(xs, self.sys.R)
                cols = [_mlp3_hess(p, f"params/orbitals_{name}/orbital_{k}", f, Jf, Hf) for k in range(ns)]
                Phi = np.stack([c[0] for c in cols], -1)                       # (W,i,k)
                G = np.stack([c[1] for c in cols], -2)                         # (W,i,k,3)
                Hs = np.stack([c[2] for c in cols], -3)                        # (W,i,k,3,3)
                caches = [c[3] for c in cols]
            else:
                from .nqmc_oracle import orbital_features
                f, _, _ = orbital_features(xs, self.sys.R)
                cols = [_mlp3(p, f"params/orbitals_{name}/orbital_{k}", f) for k in range(ns)]
                Phi = np.stack([c[0] for c in cols], -1)
                caches = [c[3] for c in cols]
            sg, la = np.linalg.slogdet(Phi)
            sign, logabs = sign * sg, logabs + la
            inv = np.linalg.inv(Phi)                                           # inv[w,k,i]
            mats.append((Phi, None, None, caches)); invs.append(inv)
            if lap:
                invT = np.swapaxes(inv, -1, -2)                                # [w,i,k]
                Jx = bf["Jx"][:, sl]                                           # (W,i,a,e,c)
                v = np.einsum("wik,wika->wia", invT, G)                        # v_i = sum_k inv[k,i] g_ik
                v_all[:, sl] = v
                gdet += np.einsum("wia,wiaec->wec", v, Jx)
                Gm = np.einsum("wiaec,wibec->wiab", Jx, Jx)                    # sum_mu d_mu x_ia d_mu x_ib
                term1 = np.einsum("wik,wikab,wiab->w", invT, Hs, Gm) + np.einsum("wia,wia->w", v, bf["lapx"][:, sl])
                dPhi = np.einsum("wika,wiaec->wecik", G, Jx)                   # d_mu Phi[i,k]
                M = np.einsum("wki,wecil->weckl", inv, dPhi)                   # Phi^-1 d_mu Phi
                term2 = np.einsum("weckl,weclk->w", M, M)
                lap_det += term1 - term2
        J, gJ, lJ, _ = self._jastrow(p, r, lap)
        logabs = logabs + J
        g = gdet + gJ if lap else None
        lother = lJ
        if self.net.use_cusp:
            C, gC, lC, _ = self._cusps(p, r, lap)
            logabs = logabs + C
            if lap:
                g = g + gC
                lother = lother + lC
        if lap:
            return sign, logabs, (g, lap_det + lother, mats, invs, p, bf, v_all)
        return sign, logabs, (None, None, mats, invs, p, bf, None)

    def grad_log_psi(self, theta, r):
        # value pass + the v_i vectors need orbital gradients -> run the derivative pass
        sign, logabs, aux = self._forward(theta, r, lap=True)
        g, lap_log, mats, invs, p, bf, v_all = aux
        Wn, N, _ = r.shape
        grads: Dict[str, np.ndarray] = {}
        for (name, sl, ns), m, inv in zip(self._spin_slices(), mats, invs):
            if m is None:
                continue
            for k in range(ns):
                _mlp3_backward(p, f"params/orbitals_{name}/orbital_{k}", m[3][k], inv[:, k, :], grads)
        # backflow network: d log|det| / d eta_p = (v_i - v_j) . (r_i - r_j)
        ii, jj, d, uh, E = bf["ii"], bf["jj"], bf["d"], bf["uh"], bf["E"]
        seed = np.einsum("wpa,wpa->wp", v_all[:, ii] - v_all[:, jj], uh * d[..., None]) * E
        _mlp3_backward(p, "params/backflow", bf["cache"], seed, grads)
        # Jastrow / cusp scalars and networks: reuse the no-backflow code on physical r
        base = Oracle(self.sys, type(self.net)(**{**self.net.__dict__, "use_backflow": False}))
        th_nb = np.concatenate([np.asarray(theta, dtype=r.dtype)[off:off + (int(np.prod(sh)) if len(sh) else 1)]
                                for path, sh, off in self.table if "/backflow/" not in path])
        O_nb = base.grad_log_psi(th_nb, r)
        O = np.zeros((Wn, self.P), dtype=r.dtype)
        nb_off = {path: off for path, _, off in base.table}
        for path, shape, off in self.table:
            n = int(np.prod(shape)) if len(shape) else 1
            if path in grads:
                O[:, off:off + n] = grads[path].reshape(Wn, n)
            else:
                O[:, off:off + n] = O_nb[:, nb_off[path]:nb_off[path] + n]
        return O
