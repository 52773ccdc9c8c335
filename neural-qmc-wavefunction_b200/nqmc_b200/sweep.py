"""Independent walker sets over several geometries (BASELINE config 4: the H4 dissociation sweep,
R = linspace(1.0, 4.0, 16), `.github/phase4-issue-body.md:168-200` scan_pes).

The reference trains one geometry after another in a Python loop.  Geometries share nothing -- each has
its own nuclei, V_nn, parameters theta and walkers -- so they are independent units: no collective, and
across GPUs geometry g simply runs on rank g % world.  Here every geometry owns a device context; one
`step()` advances all local geometries by one fused walker step (MH move, E_L, weighted gradient sums).
The geometries are dealt over ``n_streams`` side streams that fork from / join the caller's stream (``nqmc_stream_wait``):
a 16k-walker launch leaves each persistent CTA only ~7 tiles, so the tail of one geometry's kernel and the weight
staging of the next overlap, and the thread-per-walker kernels of one geometry run beside the tensor-core kernels of
another.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Sequence

import numpy as np

import ctypes as C

from . import _native, parallel
from . import random as nqrandom
from .params import ravel
from .systems.molecule import hydrogen_chain
from .wavefunction import AntisymmetricWavefunction


@dataclass
class GeometryState:
    bond_length: float
    molecule: object
    wavefunction: object
    params: dict
    ctx: object
    r: object            # device SoA [3N, W]
    logp: object         # device [W]
    e_l: object
    hdr: object
    gsum: object


class GeometrySweep:
    def __init__(self, bond_lengths: Sequence[float] = tuple(np.linspace(1.0, 4.0, 16)), n_atoms: int = 4,
                 walkers_per_geometry: int = 16384, hidden=(64, 64), use_cusp: bool = False, use_backflow: bool = False,
                 step_size: float = 0.5, seed: int = 42, n_streams: int = 8):
        self.step_size, self.seed, self.W = float(step_size), int(seed), int(walkers_per_geometry)
        self.n_streams, self.side, self.main = int(n_streams), [], 0
        rank, ws = parallel.world()
        self.local_ids = [g for g in range(len(bond_lengths)) if g % ws == rank]
        self.states: List[GeometryState] = []
        for g in self.local_ids:
            R = float(bond_lengths[g])
            mol = hydrogen_chain(n_atoms, R)
            wf = AntisymmetricWavefunction(n_atoms // 2, n_atoms // 2, hidden, use_cusp=use_cusp, use_backflow=use_backflow,
                                           molecule=mol)
            params = wf.init(nqrandom.PRNGKey(seed + g))
            ctx = wf.context(mol)
            ctx.set_params(ravel(params))
            rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, g, 0]))
            pos = np.asarray(mol.positions, np.float32)[np.arange(mol.n_electrons) % mol.n_atoms][None]
            r0 = (pos + 0.5 * rng.standard_normal((self.W, mol.n_electrons, 3))).astype(np.float32)
            r = ctx.to_soa(r0)
            _, la = ctx.log_psi(r)
            logp = ctx.upload((2.0 * la.numpy(ctx.stream)).astype(np.float32))
            self.states.append(GeometryState(R, mol, wf, params, ctx, r, logp, ctx.empty(self.W),
                                             ctx.empty(8, np.float64), ctx.empty(ctx.P, np.float64)))
        self.t = 0
        if self.n_streams > 1 and self.states:
            lib, dev = self.states[0].ctx.lib, self.states[0].ctx.device
            self.main = self.states[0].ctx.stream
            for st in self.states:
                st.ctx.synchronize()
            for _ in range(min(self.n_streams, len(self.states))):
                h = C.c_void_p()
                _native._mem_check(lib.nqmc_stream_create(dev, C.byref(h)), "nqmc_stream_create")
                self.side.append(h.value)
            for i, st in enumerate(self.states):
                st.ctx.stream = self.side[i % len(self.side)]

    def close(self):
        """Destroy the side streams (the contexts go back to the caller's stream)."""
        if self.side and self.states:
            self.synchronize()
            c = self.states[0].ctx
            for st in self.states:
                st.ctx.stream = self.main
            for s_ in self.side:
                c.lib.nqmc_stream_destroy(c.device, s_)
            self.side = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _fork(self):
        if self.side:
            c = self.states[0].ctx
            for s in self.side:
                _native._mem_check(c.lib.nqmc_stream_wait(c.device, s, self.main), "nqmc_stream_wait")

    def _join(self):
        if self.side:
            c = self.states[0].ctx
            for s in self.side:
                _native._mem_check(c.lib.nqmc_stream_wait(c.device, self.main, s), "nqmc_stream_wait")

    def burn_in(self, n_steps: int):
        self._fork()
        for t in range(n_steps):
            for st, g in zip(self.states, self.local_ids):
                st.ctx.mh_step(st.r, st.logp, self.step_size, seed=self.seed + 1000 * g, step=t)
        self._join()

    def step(self):
        """One fused walker step for every local geometry; results stay on the device until ``results()``."""
        self.t += 1
        self._fork()
        for st, g in zip(self.states, self.local_ids):
            st.ctx.walker_step(st.r, st.logp, self.step_size, st.e_l, st.hdr, st.gsum, seed=self.seed + 1000 * g,
                               step=(1 << 20) + self.t)
        self._join()

    def synchronize(self):
        for st in self.states:
            st.ctx.synchronize()

    def results(self):
        out = []
        for st in self.states:
            grad, mean, std = parallel.combine_gradient(st.hdr.numpy(st.ctx.stream), st.gsum.numpy(st.ctx.stream))
            out.append({"bond_length": st.bond_length, "energy": mean, "std": std, "gradient": grad,
                        "v_nn": st.molecule.nuclear_repulsion_energy()})
        return out
