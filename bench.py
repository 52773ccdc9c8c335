#!/usr/bin/env python
"""bench.py -- walker-steps/s of the nqmc VMC walker loop on B200 (BASELINE.json metric).

One "step" = one walker-step for every walker of the batch: one Metropolis-Hastings
proposal + accept/reject (value-only psi), the local energy E_L at the resulting position (forward
Laplacian), and the weighted log-derivative gradient accumulation sum_w (E_w - <E>) O_w, with the
per-iteration all-reduces of the energy header and the gradient vector when N > 1.

Workload at every N: BASELINE configs[1] -- H4 linear chain (2 up, 2 down, R = 1.8 bohr), neural
Slater-Jastrow with (64,64) orbital networks (P = 19,976), 65,536 walkers PER GPU (weak scaling),
synthetic walkers (nuclei + 0.5 N(0,1), then >= 100 device MH burn-in steps), LeCun-normal weights.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference ...                     (CPU restatement of the reference algorithm)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "neural-qmc-wavefunction_b200"))

N_ATOMS, BOND, HIDDEN, SIGMA = 4, 1.8, 64, 0.5          # step_size 0.5 for H4 (phase4-issue-body.md:139)
WALKERS_PER_GPU = 65536
SEED = 42                                               # scripts/train_h2.py:63


def algorithmic_flops_per_walker_step(n_atoms, hidden, n_elec, backflow=False):
    """SURVEY 8(d) / BASELINE.md section 3: 2 m E (3 + c)."""
    F = 3 + 2 * n_atoms
    m = F * hidden + hidden * hidden + hidden
    E = 2 * (n_elec // 2) ** 2
    c = 10 if backflow else 5
    return 2.0 * m * E * (3 + c), 2.0 * m * E


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  The timed region lasts tens of milliseconds,
    so NVML is polled in-process every ~1 ms (nvidia-smi -lms cannot sample that fast); falls back to one
    nvidia-smi query when NVML is unavailable."""

    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.max_mhz = index, [], 0, None
        self._stop = threading.Event()
        self._thread = None
        self._h = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._thread = threading.Thread(target=self._poll, daemon=True)
            self._thread.start()
        except Exception:
            self._h = None

    def _poll(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            except Exception:
                pass
            time.sleep(0.001)

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
        if self._h is None:                       # fallback: one nvidia-smi sample (still under load: called before sync)
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": a, "sm_max_mhz": b, "reasons": [], "samples": 1, "source": "nvidia-smi"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm), "source": "nvml polled every ~1 ms during the timed steps"}


def as_torch(x):
    """DeviceArray -> torch CUDA tensor over the same memory (bench plumbing only)."""
    import torch
    if isinstance(x, torch.Tensor):
        return x
    t = torch.as_tensor(x, device=torch.device("cuda", x.device))
    t._keep = x
    return t


def sustained_run(ctx, one_step, Wl, flops_step, local, peak3, tf32_rec, seconds=2.5):
    """>= 2 s of back-to-back fused steps (no L2 flush: the 3 MB of positions is the only reused data) with its own
    clock record: what the step delivers once the 1 kW power cap, not the burst clock, sets the pace."""
    import torch
    cs = ClockSampler(local)
    cs.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 0
    while time.perf_counter() - t0 < seconds:
        for _ in range(100):
            one_step()
        n += 100
        torch.cuda.synchronize()
    e1.record()
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    clocks = cs.stop()
    tfl = flops_step * Wl * n / (ms * 1e-3) / 1e12
    return {"value": Wl * n / (ms * 1e-3), "unit": "walker-steps/s", "steps": n, "seconds": ms * 1e-3, "ms_per_step": ms / n,
            "whole_step_tflops": tfl, "frac_of_burst_peak": tfl / peak3,
            "frac_of_sustained_peak": tfl / (tf32_rec["sustained_tflops"] / 3.0), "clocks": clocks,
            "note": "back-to-back, L2 not flushed, host synchronises every 100 steps"}


def vmc_iteration_rate(W, iters=20):
    """Iterations/s of the drop-in VMCOptimizer.step itself (vmc.py:217-282) at the benchmark size: walkers resident
    on the device, 10 value-only burn-in moves (vmc.py:247) + the fused sampled step, fused clip + Adam on the device,
    four metrics read back per iteration (the reference's float() calls, vmc.py:275-280)."""
    from nqmc_b200 import random as nqrandom
    from nqmc_b200.optimizer import VMCOptimizer
    mol, wf, params, theta = make_problem()
    opt = VMCOptimizer(learning_rate=1e-4, n_samples=1, n_chains=W, n_burn=10, step_size=SIGMA, device_update=True)
    key = nqrandom.PRNGKey(SEED)
    key, k0 = nqrandom.split(key)
    st = opt.init(k0, wf, params, mol)
    for _ in range(3):
        key, ks = nqrandom.split(key)
        st, m = opt.step(ks, st, wf, mol)
    t0 = time.perf_counter()
    for _ in range(iters):
        key, ks = nqrandom.split(key)
        st, m = opt.step(ks, st, wf, mol)
    dt = (time.perf_counter() - t0) / iters
    return {"iterations_per_s": 1.0 / dt, "ms_per_iteration": dt * 1e3, "sampled_walker_steps_per_s": W / dt,
            "mh_moves_per_s": 11 * W / dt, "n_chains": W, "n_samples": 1, "n_burn": 10, "device_update": True,
            "last_metrics": m, "timing": "host wall clock around VMCOptimizer.step (it synchronises to return its metrics)"}


def sr_syrk_rate(ctx, peak3):
    """S = O^T O / n + reg I for P parameters x n samples on the tcgen05 3xTF32 SYRK (nqmc_gram, TMA-fed),
    device time by CUDA events.  P = 4,673 is config 1's SimpleWavefunction (64, 64); 19,976 the H4 ansatz of the headline."""
    import torch
    dev = torch.device("cuda", ctx.device)
    out = []
    for P, n in ((4673, 65536), (19976, 16384)):
        g = torch.Generator(device=dev).manual_seed(1)
        Ot = torch.randn((P, n), device=dev, dtype=torch.float32, generator=g)
        S = torch.empty((P, P), device=dev, dtype=torch.float32)
        torch.cuda.synchronize()

        def run():
            rc = ctx.lib.nqmc_gram(ctx.h, Ot.data_ptr(), Ot.data_ptr(), P, P, n, n, n, 1.0 / n, 1e-3, 1, S.data_ptr(), P, ctx.stream)
            assert rc == 0, rc
        run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            run()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        tiles = sum(1 for tm in range((P + 127) // 128) for tn in range((P + 255) // 256) if tn * 256 + 256 > tm * 128)
        alg, exe = n * P * (P + 1.0), tiles * 128 * 256 * 2.0 * n
        out.append({"P": P, "n_samples": n, "ms": ms, "algorithmic_tflops": alg / ms / 1e9, "executed_tflops": exe / ms / 1e9,
                    "frac_of_tf32_over_3_executed": exe / ms / 1e9 / peak3, "frac_of_tf32_over_3_algorithmic": alg / ms / 1e9 / peak3})
        del Ot, S
        torch.cuda.empty_cache()
    return {"kernel": "gram_kernel (nqmc_gram.cu): TMA SWIZZLE_128B -> smem ring -> tf32 hi/lo split -> tcgen05.mma M128 N256 K8, "
                      "upper-triangle tiles only, fp32 flush every 256 samples", "cases": out,
            "note": "algorithmic = n P (P+1) flops (upper triangle); executed = tiles touching it"}


def sr_iteration_rate(iters=5):
    """One full Stochastic-Reconfiguration VMC iteration through the drop-in API (VMCOptimizer(optimizer='sr')): config 1's
    H2 SimpleWavefunction (64, 64), P = 4,673, 8,192 chains x 8 samples = 65,536 samples per iteration -- sampling, E_L,
    per-walker O on the device, tcgen05 SYRK, CG solve, theta update."""
    from nqmc_b200 import random as nqrandom
    from nqmc_b200.optimizer import VMCOptimizer
    from nqmc_b200.systems.molecule import hydrogen_molecule
    from nqmc_b200.wavefunction import SimpleWavefunction
    mol = hydrogen_molecule(1.4)
    wf = SimpleWavefunction(hidden_dims=(64, 64), envelope_decay=1.0, nuclear_positions=mol.positions)
    key = nqrandom.PRNGKey(SEED)
    key, k0, k1 = nqrandom.split(key, 3)
    params = wf.init(k0, np.zeros((2, 3)))
    opt = VMCOptimizer(learning_rate=0.05, n_samples=8, n_chains=8192, n_burn=50, step_size=0.3, optimizer="sr",
                       sr_regularization=1e-3, sr_max_iter=100, sr_tol=1e-5)
    st = opt.init(k1, wf, params, mol)
    for _ in range(2):
        key, ks = nqrandom.split(key)
        st, m = opt.step(ks, st, wf, mol)
    t0 = time.perf_counter()
    for _ in range(iters):
        key, ks = nqrandom.split(key)
        st, m = opt.step(ks, st, wf, mol)
    dt = (time.perf_counter() - t0) / iters
    return {"iterations_per_s": 1.0 / dt, "ms_per_iteration": dt * 1e3, "P": 4673, "samples_per_iteration": 65536,
            "last_metrics": m, "timing": "host wall clock around VMCOptimizer.step"}


def sr_iteration_rate_h4(iters=3):
    """The same for the benchmark wavefunction (H4, four 11-64-64-1 orbital networks, P = 19,976): 8,192 chains x 8 samples =
    65,536 samples per iteration; per-walker O from nqmc_pergrad.cu (5.2 GB), S = 1.6 GB."""
    from nqmc_b200 import random as nqrandom
    from nqmc_b200.optimizer import VMCOptimizer
    from nqmc_b200.systems.molecule import hydrogen_chain
    from nqmc_b200.wavefunction import AntisymmetricWavefunction
    mol = hydrogen_chain(4, 1.8)
    wf = AntisymmetricWavefunction(2, 2, (64, 64), molecule=mol)
    key = nqrandom.PRNGKey(SEED)
    key, k0, k1 = nqrandom.split(key, 3)
    params = wf.init(k0)
    opt = VMCOptimizer(learning_rate=0.05, n_samples=8, n_chains=8192, n_burn=50, step_size=0.5, optimizer="sr",
                       sr_regularization=1e-3, sr_max_iter=100, sr_tol=1e-5)
    st = opt.init(k1, wf, params, mol)
    key, ks = nqrandom.split(key)
    st, m = opt.step(ks, st, wf, mol)
    t0 = time.perf_counter()
    for _ in range(iters):
        key, ks = nqrandom.split(key)
        st, m = opt.step(ks, st, wf, mol)
    dt = (time.perf_counter() - t0) / iters
    return {"iterations_per_s": 1.0 / dt, "ms_per_iteration": dt * 1e3, "P": 19976, "samples_per_iteration": 65536,
            "last_metrics": m, "timing": "host wall clock around VMCOptimizer.step"}


def other_configs_sharded(flush, rank, world):
    """configs[2] and configs[4] -- the two BASELINE configs quoted on 8 GPUs -- weak-scaled over `world` ranks at their
    per-GPU sizes (262,144 / 8 and 1,048,576 / 8 walkers per GPU).  Collective: every rank calls this; device time, max
    over ranks.  The peak fractions are added by rank 0 (add_fracs)."""
    out = {}
    for name, args_ in (("configs[2] H6 +backflow +cusp, orbital MLP 15-128-128-1 x6", (6, 128, 32768, True, True)),
                        ("configs[4] H10, orbital MLP 23-128-128-1 x10", (10, 128, 131072, False, False))):
        out[name] = time_fused_step(*args_, steps=5, warm=2, flush=flush, rank=rank, world=world)
    return out


def other_configs(flush, peak3, ffma_peak):
    """The BASELINE configs that are not the headline, on ONE GPU at their per-GPU sizes (configs[2] and [4] name 8
    GPUs: 262,144 / 8 and 1,048,576 / 8 walkers), same event timing and L2 flush as the headline."""
    import torch
    out = {}
    for name, args_ in (("configs[2] H6 +backflow +cusp, orbital MLP 15-128-128-1 x6", (6, 128, 32768, True, True)),
                        ("configs[4] H10, orbital MLP 23-128-128-1 x10", (10, 128, 131072, False, False)),
                        ("H6 without backflow/cusp, 128-wide (tensor-core forward)", (6, 128, 32768, False, False))):
        try:
            d = time_fused_step(*args_, steps=5, warm=2, flush=flush)
            d["frac_of_tf32_over_3"] = d["achieved_tflops"] / peak3
            d["frac_of_ffma_peak"] = d["achieved_tflops"] / ffma_peak
            d["kernels"] = ("tcgen05 3xTF32: value + 5-channel forward (nqmc_orbital_fwd_pc.cu; backflow: G-weighted Laplacian "
                            "channel) and the reverse pass (nqmc_orbital_bwd128.cu: value sweep, D1 kernel, TMA-fed "
                            "weight-gradient Gram kernel)"
                            + ("; FP32: backflow / cusp per-walker kernels" if args_[4] else ""))
            out[name] = d
        except Exception as e:      # report, do not hide
            out[name] = {"error": repr(e)}
    try:
        from nqmc_b200.sweep import GeometrySweep
        sw = GeometrySweep(bond_lengths=np.linspace(1.0, 4.0, 16), walkers_per_geometry=16384, hidden=(64, 64),
                           step_size=SIGMA, seed=SEED)
        sw.burn_in(30)
        for _ in range(2):
            sw.step()
        sw.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
        for a, b in ev:
            flush.fill_(1.0)
            a.record()
            sw.step()
            b.record()
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        fl, _ = algorithmic_flops_per_walker_step(4, 64, 4)
        tot = 16 * 16384
        out["configs[3] H4 dissociation sweep, 16 geometries R=1.0..4.0 x 16,384 walkers (independent sets, no collective)"] = {
            "walkers_per_gpu": tot, "ms_per_step": ms, "value_per_gpu": tot / (ms * 1e-3), "unit": "walker-steps/s",
            "achieved_tflops": fl * tot / (ms * 1e-3) / 1e12, "frac_of_tf32_over_3": fl * tot / (ms * 1e-3) / 1e12 / peak3,
            "energies": [round(o["energy"], 4) for o in sw.results()][:4],
            "kernels": "all three orbital kernels tcgen05 3xTF32; 16 contexts x 9 launches dealt over 8 streams that fork from / join the timed stream"}
    except Exception as e:
        out["configs[3] sweep"] = {"error": repr(e)}
    return out


def make_problem():
    from nqmc_b200.systems import hydrogen_chain
    from nqmc_b200.wavefunction import AntisymmetricWavefunction
    from nqmc_b200 import random as nqrandom
    from nqmc_b200.params import ravel
    mol = hydrogen_chain(N_ATOMS, BOND)
    wf = AntisymmetricWavefunction(n_up=2, n_down=2, orbital_hidden_dims=(HIDDEN, HIDDEN), molecule=mol)
    params = wf.init(nqrandom.PRNGKey(SEED))
    return mol, wf, params, ravel(params)


def initial_walkers(mol, n, lo, seed=SEED):
    """electron i at nucleus i % A + 0.5 N(0,1) (optimizer/vmc.py:199-208), keyed by global walker id."""
    rng = np.random.Generator(np.random.Philox(key=seed, counter=[0, 0, 0, lo]))
    R = np.asarray(mol.positions, dtype=np.float32)
    init = np.stack([R[i % mol.n_atoms] for i in range(mol.n_electrons)], axis=0)
    return (init[None] + 0.5 * rng.standard_normal((n, mol.n_electrons, 3))).astype(np.float32)


# =================================================================================================
def run_ours(args):
    import torch
    import torch.distributed as dist
    from nqmc_b200 import parallel

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    Wl = args.walkers_per_gpu
    lo = rank * Wl
    mol, wf, params, theta = make_problem()
    ctx = wf.context(mol)
    ctx.set_params(theta)
    ctx.reserve(Wl)
    # N > 1: hdr / gsum are summed over ranks inside nqmc_walker_step by the peer-memory all-reduce (NVLink);
    # NQMC_COMM=nccl keeps the two NCCL all-reduces around the rank-local halves instead (A/B runs)
    peer = world > 1 and os.environ.get("NQMC_COMM", "peer") != "nccl" and parallel.attach_peer_memory(ctx)
    r = ctx.to_soa(initial_walkers(mol, Wl, lo))
    _, la = ctx.log_psi(r)
    logp = (2.0 * as_torch(la)).contiguous()
    e_l = torch.empty(Wl, dtype=torch.float32, device=dev)
    hdr = torch.empty(8, dtype=torch.float64, device=dev)
    gsum = torch.empty(ctx.P, dtype=torch.float64, device=dev)
    nacc = torch.zeros(Wl, dtype=torch.float32, device=dev)
    for t in range(args.burn):                                 # equilibrate: walkers ~ |psi|^2, away from nodes
        ctx.mh_step(r, logp, SIGMA, seed=SEED, step=t, walker_id0=lo)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    step_no = [1 << 20]

    def one_step():
        s = step_no[0]
        step_no[0] += 1
        if peer:
            ctx.walker_step(r, logp, SIGMA, e_l, hdr, gsum, seed=SEED, step=s, walker_id0=lo, n_accepted=nacc)
            return
        ctx.walker_step_a(r, logp, SIGMA, e_l, hdr, seed=SEED, step=s, walker_id0=lo, n_accepted=nacc)
        if world > 1:
            dist.all_reduce(hdr)
        ctx.walker_step_b(r, e_l, hdr, gsum)
        if world > 1:
            dist.all_reduce(gsum)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        one_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.fill_(1.0)                                       # L2 flush between timed steps (not timed)
        a.record()
        one_step()
        b.record()
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in ev)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel durations for the roofline: a separate, untimed pass (the event records between launches would
    # otherwise sit inside the timed region)
    ctx.profile(True)
    for _ in range(min(args.steps, 20)):
        flush.fill_(1.0)
        one_step()
    barrier()
    prof = {k: ctx.profile_read(i) for i, k in enumerate(["eval1", "eval5", "bwd", "assemble", "accept"])}
    ctx.profile(False)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    h = hdr.cpu().numpy()

    # ---- SURVEY 8(d) secondary figure: the reference's inter-iteration burn-in (vmc.py:247, n_burn=10 value-only MH
    # steps before every sampled step), counting only the sampled steps
    def burn10_iter(k):
        for j in range(10):
            ctx.mh_step(r, logp, SIGMA, seed=SEED + 1, step=(1 << 23) + 10 * k + j, walker_id0=lo)
        one_step()
    nb = max(args.steps // 5, 4)
    for k in range(2):
        burn10_iter(k)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record()
    for k in range(nb):
        burn10_iter(2 + k)
    b1.record()
    barrier()
    tb = torch.tensor([b0.elapsed_time(b1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tb, op=dist.ReduceOp.MAX)
    burn10_value = world * Wl * nb / (float(tb.item()) * 1e-3)

    # ---- e2e: host buffers in pinned memory, copies inside the timed region -----------------------
    nn = 3 * mol.n_electrons
    r_host = torch.empty((Wl, mol.n_electrons, 3), dtype=torch.float32).pin_memory()
    r_host.copy_(torch.from_numpy(ctx.to_aos(r).numpy()))
    logp_host = torch.empty(Wl, dtype=torch.float32).pin_memory(); logp_host.copy_(logp.cpu())
    e_host = torch.empty(Wl, dtype=torch.float32).pin_memory()
    hdr_host = torch.empty(8, dtype=torch.float64).pin_memory()
    g_host = torch.empty(ctx.P, dtype=torch.float64).pin_memory()
    h2d = Wl * nn * 4 + Wl * 4
    d2h = Wl * nn * 4 + Wl * 4 + Wl * 4 + 8 * 8 + ctx.P * 8

    rh, lh, eh, hh, gh = (x.numpy() for x in (r_host, logp_host, e_host, hdr_host, g_host))
    pipelined = [os.environ.get("NQMC_E2E_SYNC", "0") != "1"]

    def e2e_step(s, last=False):
        if (world == 1 or peer) and pipelined[0]:              # the C-ABI host entry points, pipelined form: every step's
            # inputs go H2D and its results D2H inside the loop; the upload of step s+1's walkers (available once phase A
            # of step s has come home) overlaps the reverse pass of step s
            ctx.walker_step_host_async(rh, lh, SIGMA, eh, hh, gh, seed=SEED, step=s, walker_id0=lo)
            ctx.walker_step_host_wait(ctx.WAIT_PHASE_A)
            if not last:
                ctx.walker_step_host_upload(rh, lh)
            ctx.walker_step_host_wait(ctx.WAIT_ALL)
        elif world == 1 or peer:                               # synchronous host entry point
            ctx.walker_step_host(rh, lh, SIGMA, eh, hh, gh, seed=SEED, step=s, walker_id0=lo)
        else:                                                  # same copies, phases split around the all-reduces
            ra = r_host.to(dev, non_blocking=True)
            lp = logp_host.to(dev, non_blocking=True)
            rs = ctx.to_soa(ra)
            ctx.walker_step_a(rs, lp, SIGMA, e_l, hdr, seed=SEED, step=s, walker_id0=lo)
            dist.all_reduce(hdr)
            ctx.walker_step_b(rs, e_l, hdr, gsum)
            dist.all_reduce(gsum)
            r_host.copy_(as_torch(ctx.to_aos(rs)), non_blocking=True)
            logp_host.copy_(lp, non_blocking=True)
            e_host.copy_(e_l, non_blocking=True)
            hdr_host.copy_(hdr, non_blocking=True)
            g_host.copy_(gsum, non_blocking=True)
            torch.cuda.synchronize()

    def e2e_run(base):
        for i in range(3):
            e2e_step(base + i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(base + 64 + i, last=(i == args.steps - 1))
        dt = time.perf_counter() - t0                          # this rank's own finish (every step ends synchronised);
        barrier()                                              # the max over ranks is taken below
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    e2e_s = e2e_run(1 << 21)
    e2e_sync_s = None
    if pipelined[0] and (world == 1 or peer):                  # the synchronous entry point, for comparison
        pipelined[0] = False
        e2e_sync_s = e2e_run(1 << 22)
        pipelined[0] = True

    sharded = None
    if world > 1 and peer and not args.no_extras:                # configs[2] / configs[4] over the same ranks
        try:
            sharded = other_configs_sharded(flush, rank, world)
        except Exception as e:      # report, do not hide
            sharded = {"error": repr(e)}

    out = None
    if rank == 0:
        flops_step, flops_value = algorithmic_flops_per_walker_step(N_ATOMS, HIDDEN, mol.n_electrons)
        pk = peaks()
        ffma_peak = ctx.ffma_peak(5)
        tf32_rec = measure_tf32_peak(dev, local)
        tf32_peak = tf32_rec["burst_tflops"]
        use_tc = os.environ.get("NQMC_TC", "1") != "0"
        ms5, n5 = prof["eval5"]
        # dominant kernel: the 5-channel orbital forward (value, gradient, Laplacian): 5 * 2 m E flop per walker
        flops_launch = 5.0 * flops_value * Wl
        ach = flops_launch / (ms5 / max(n5, 1) * 1e-3) / 1e12 if n5 else None
        F_in = 3 + 2 * N_ATOMS
        frac_l2 = HIDDEN * HIDDEN / float(F_in * HIDDEN + HIDDEN * HIDDEN + HIDDEN)      # share of the MACs in the H x H layer
        peak3 = tf32_peak / 3.0                                   # 3xTF32: three tensor-core products per fp32-grade product
        # time floor of the kernel: H x H layer at the 3xTF32 tensor rate + input/output layers at the FP32 FFMA rate
        t_floor = flops_launch * frac_l2 / (peak3 * 1e12) + flops_launch * (1 - frac_l2) / (ffma_peak * 1e12)
        value = world * Wl * args.steps / (total_ms * 1e-3)
        whole_tflops = flops_step * Wl / (total_ms / args.steps * 1e-3) / 1e12
        out = {
            "metric": "walker-steps/s", "value": value, "unit": "walker-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "H4 chain R=1.8 (2up/2dn) neural Slater-Jastrow, orbital MLP 11-64-64-1 x4, P=19976",
                       "walkers_per_gpu": Wl, "global_walkers": world * Wl, "step_size": SIGMA,
                       "rng": "device Philox4x32-10", "l2": "256 MiB flush write between timed steps",
                       "parallelism": f"walkers sharded over {world} GPU(s); all-reduce hdr(8 f64)+grad(P f64) per step",
                       "collective": ("none" if world == 1 else "peer-memory one-shot all-reduce inside nqmc_walker_step (NVLink)"
                                      if peer else "NCCL all_reduce x2"),
                       "algorithmic_flop_per_walker_step": flops_step},
            "with_burn10": {"value": burn10_value, "unit": "sampled walker-steps/s",
                            "note": "10 value-only MH steps before every sampled step (reference n_burn between iterations, "
                                    "vmc.py:247); only sampled steps counted; L2 not flushed"},
            "clocks": clocks,
            "e2e": {"value": world * Wl * args.steps / e2e_s, "unit": "walker-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h,
                    "path": ("nqmc_walker_step_host_async / _wait / _upload (C ABI, pinned host buffers; the H2D of step s+1's "
                             "walkers overlaps the reverse pass of step s)" if e2e_sync_s is not None else
                             "nqmc_walker_step_host (C ABI, pinned host buffers)") if (world == 1 or peer)
                    else "pinned host buffers -> nqmc_walker_step_a/b + NCCL all-reduce -> host",
                    "synchronous_value": (world * Wl * args.steps / e2e_sync_s) if e2e_sync_s else None,
                    "synchronous_path": "nqmc_walker_step_host (one blocking call per step)" if e2e_sync_s else None},
            "gpu_launches": launches,
            "roofline": ({"bound": "tensor", "kernel": "orb_lap_pc_kernel<4> (orbital forward value+grad+Laplacian; H x H layer = tcgen05 3xTF32 with A in TMEM, input layer = FFMA)",
                          "achieved": ach, "peak": peak3, "unit": "TFLOP/s", "frac": (ach / peak3) if ach else None,
                          "peak_source": "cuBLAS TF32 GEMM measured live in this run (50 reps at 4096^3 and 8192^3, max = burst), "
                                         "divided by 3 (3xTF32 split); MEASURED_PEAKS.json has only a bf16 figure",
                          "tf32_dense_tflops_measured": tf32_peak, "tf32_peaks": tf32_rec,
                          # the same kernel counted three ways
                          "frac_algorithmic": (ach / peak3) if ach else None,
                          "frac_tensor_only": (ach * frac_l2 / peak3) if ach else None,
                          "frac_tensor_only_note": "only the H x H layer's flops (the part executed on tensor cores) over the "
                                                   "3xTF32 peak; the input and output layers run on the FP32 pipe",
                          "whole_step": {"achieved": whole_tflops, "frac": whole_tflops / peak3,
                                         "frac_vs_sustained_peak": whole_tflops / (tf32_rec["sustained_tflops"] / 3.0),
                                         "note": "algorithmic flops of the whole fused walker step (SURVEY 8(d), 2mE(3+c)) over "
                                                 "its device time, against the same tensor denominator"},
                          "traffic": 3.2735e6, "traffic_note": "bytes per launch: dram__bytes_read.sum + dram__bytes_write.sum "
                          "from the ncu --set full capture in profiles/r2s_ncu_full_metrics.txt (3.27 MB read = the walker positions, "
                          "0 written: the 10.5 MB of orbital channels stay in the 126 MB L2 for the next kernel)",
                          "frac_of_mixed_floor": (t_floor / (ms5 / max(n5, 1) * 1e-3)) if n5 else None,
                          "mixed_floor_note": "floor = HxH-layer flops / (tf32/3) + other-layer flops / measured FFMA peak",
                          "frac_vs_bf16_peak": (ach / pk["bf16_tflops"]) if (ach and pk.get("bf16_tflops")) else None}
                         if use_tc else
                         {"bound": "fp32_ffma", "kernel": "orb_eval_kernel<64,5,4,256> (orbital forward: value+grad+Laplacian)",
                          "achieved": ach, "peak": ffma_peak, "unit": "TFLOP/s", "frac": (ach / ffma_peak) if ach else None,
                          "peak_source": "measured live: nqmc_ffma_peak (register FFMA chains)"}),
            "roofline_context": {"ffma_peak_tflops_measured": ffma_peak, "peak_nominal_148x128x2x1.965GHz": 74.4,
                         "bf16_tensor_peak_measured": pk.get("bf16_tflops"),
                         "hbm_gbs_measured": pk.get("hbm_gbs"),
                         "whole_step_achieved_tflops": flops_step * Wl / (total_ms / args.steps * 1e-3) / 1e12,
                         "kernel_ms": {k: (v[0] / max(v[1], 1)) for k, v in prof.items()},
                         # algorithmic TFLOP/s of the three orbital kernels (value 2mE, Laplacian 5 x 2mE,
                         # parameter gradient 2 x 2mE per walker) and their fraction of the 3xTF32 tensor peak
                         "kernel_tflops": {k: (mult * flops_value * Wl / (prof[k][0] / max(prof[k][1], 1) * 1e-3) / 1e12)
                                           for k, mult in (("eval1", 1.0), ("eval5", 5.0), ("bwd", 2.0)) if prof[k][1]},
                         "kernel_frac_of_tf32_over_3": {k: (mult * flops_value * Wl / (prof[k][0] / max(prof[k][1], 1) * 1e-3) / 1e12) / peak3
                                                        for k, mult in (("eval1", 1.0), ("eval5", 5.0), ("bwd", 2.0)) if prof[k][1]},
                         "kernel_launches": {k: v[1] for k, v in prof.items()}},
            "sanity": {"mean_E_L": h[1] / max(h[0], 1), "n_finite": h[0], "n_bad": h[4], "acceptance": h[3] / (world * Wl)},
        }
        if world == 1 and not args.no_extras:
            out["sustained"] = sustained_run(ctx, one_step, Wl, flops_step, local, peak3, tf32_rec)
            out["vmc_iteration"] = vmc_iteration_rate(Wl)
            out["other_configs"] = other_configs(flush, peak3, ffma_peak)
            try:
                out["sr_syrk"] = sr_syrk_rate(ctx, peak3)
            except Exception as e:      # report, do not hide
                out["sr_syrk"] = {"error": repr(e)}
            try:
                out["sr_iteration"] = sr_iteration_rate()
            except Exception as e:
                out["sr_iteration"] = {"error": repr(e)}
            try:
                out["sr_iteration_h4"] = sr_iteration_rate_h4()
            except Exception as e:
                out["sr_iteration_h4"] = {"error": repr(e)}
        if sharded is not None:
            for d in sharded.values():
                if isinstance(d, dict) and "achieved_tflops" in d:
                    d["frac_of_tf32_over_3"] = d["achieved_tflops"] / peak3
                    d["note"] = "per-GPU figures (achieved_tflops, value_per_gpu) and the whole-job value over n_gpus ranks"
            out["other_configs"] = sharded
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"] = cpu_baseline(args.cpu_walkers, steps=2, warmup=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def measure_tf32_peak(dev, local=0, sizes=(4096, 8192), reps=50):
    """Dense TF32 tensor-core throughput of this GPU (cuBLAS through torch.matmul): `reps` timed GEMMs at each of two
    sizes, the maximum is the burst peak; also a back-to-back loop of >= 1.5 s for the sustained figure, with the
    clocks seen.  MEASURED_PEAKS.json carries only a bf16 number, so the 3xTF32 roofline denominator is measured here
    the way the driver measures its bf16 entry.  Returns a dict (also written to gpurun_out/peaks_tf32.json)."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    rec = {"how": f"torch.matmul fp32 with allow_tf32 (cuBLAS), {reps} timed reps per size after 3 warm-ups, CUDA events; "
                  "burst = max over reps and sizes; sustained = back-to-back for >= 1.5 s", "per_size": {}}
    best = 0.0
    for n in sizes:
        a = torch.randn(n, n, device=dev, dtype=torch.float32)
        b = torch.randn(n, n, device=dev, dtype=torch.float32)
        vals = []
        for i in range(reps + 3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            e1.synchronize()
            if i >= 3:
                vals.append(2.0 * n ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        rec["per_size"][str(n)] = {"max": max(vals), "median": float(np.median(vals)), "min": min(vals)}
        best = max(best, max(vals))
        if n == sizes[-1]:
            cs = ClockSampler(local)
            cs.start()
            t0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            k = 0
            while time.perf_counter() - t0 < 1.5:
                for _ in range(20):
                    torch.matmul(a, b)
                k += 20
                torch.cuda.synchronize()
            e1.record()
            e1.synchronize()
            rec["sustained_tflops"] = 2.0 * n ** 3 * k / (e0.elapsed_time(e1) * 1e-3) / 1e12
            rec["sustained_clocks"] = cs.stop()
        del a, b
    torch.backends.cuda.matmul.allow_tf32 = old
    rec["burst_tflops"] = best
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(rec, open(os.path.join(ROOT, "gpurun_out", "peaks_tf32.json"), "w"), indent=1)
    except Exception:
        pass
    return rec


def time_fused_step(n_atoms, hidden, W, cusp, bf, steps, warm, flush, sigma=0.3, rank=0, world=1):
    """Device time of nqmc_walker_step for one BASELINE config at its per-GPU size (events on the launching stream,
    L2 flushed between timed steps).  world > 1: every rank runs its own W walkers (global ids rank*W ...), hdr and gsum
    are summed over ranks inside the step by the peer-memory all-reduce, the time is the max over ranks.  -> dict"""
    import torch
    import torch.distributed as dist
    from nqmc_b200 import parallel
    from nqmc_b200.systems import hydrogen_chain
    from nqmc_b200.wavefunction import AntisymmetricWavefunction
    from nqmc_b200 import random as nqrandom
    from nqmc_b200.params import ravel
    mol = hydrogen_chain(n_atoms, BOND)
    wf = AntisymmetricWavefunction(n_atoms // 2, n_atoms // 2, (hidden, hidden), use_cusp=cusp, use_backflow=bf, molecule=mol)
    theta = ravel(wf.init(nqrandom.PRNGKey(SEED)))
    ctx = wf.context(mol)
    ctx.set_params(theta)
    lo = rank * W
    if world > 1 and not parallel.attach_peer_memory(ctx):
        raise RuntimeError("peer-memory communicator could not be attached")
    r = ctx.to_soa(initial_walkers(mol, W, lo))
    _, la = ctx.log_psi(r)
    logp = ctx.upload((2.0 * la.numpy()).astype(np.float32))
    e_l, hdr, gsum = ctx.empty(W), ctx.empty(8, np.float64), ctx.empty(ctx.P, np.float64)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for t in range(30):
        ctx.mh_step(r, logp, sigma, seed=SEED, step=t, walker_id0=lo)
    for t in range(warm):
        ctx.walker_step(r, logp, sigma, e_l, hdr, gsum, seed=SEED, step=100 + t, walker_id0=lo)
    sync()
    l0 = ctx.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for t, (a, b) in enumerate(ev):
        flush.fill_(1.0)
        a.record()
        ctx.walker_step(r, logp, sigma, e_l, hdr, gsum, seed=SEED, step=1000 + t, walker_id0=lo)
        b.record()
    sync()
    ms = sum(a.elapsed_time(b) for a, b in ev) / steps
    if world > 1:
        tm = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ms = float(tm.item())
    fl, _ = algorithmic_flops_per_walker_step(n_atoms, hidden, n_atoms, backflow=bf)
    h = hdr.numpy()
    out = {"walkers_per_gpu": W, "P": ctx.P, "ms_per_step": ms, "value_per_gpu": W / (ms * 1e-3), "unit": "walker-steps/s",
           "n_gpus": world, "global_walkers": world * W, "value": world * W / (ms * 1e-3),
           "algorithmic_flop_per_walker_step": fl, "achieved_tflops": fl * W / (ms * 1e-3) / 1e12,
           "launches_per_step": (ctx.launch_count() - l0) / steps, "mean_E_L": h[1] / max(h[0], 1), "n_bad": h[4]}
    sync()                                                     # no rank frees its inbox while a peer may still push
    del ctx
    return out


# =================================================================================================
def cpu_baseline(n_walkers, steps, warmup):
    """The reference's algorithm (vmap + dense-Hessian kinetic energy + vmap(grad wrt params)) restated with
    torch.func (oracle/nqmc_autodiff.py), float32, on all host cores, on a bounded sample of the workload."""
    import torch
    from oracle.layout import OracleNet, hydrogen_chain_system, init_theta
    from oracle.nqmc_autodiff import AutodiffOracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    s = hydrogen_chain_system(N_ATOMS, BOND)
    net = OracleNet(kind=1, hidden=(HIDDEN, HIDDEN))
    th = init_theta(s, net, SEED).astype(np.float32)
    a = AutodiffOracle(s, net, dtype=torch.float32)
    rng = np.random.default_rng(SEED)
    r = (s.R[np.arange(s.n_elec) % s.n_atoms][None] + 0.5 * rng.standard_normal((n_walkers, s.n_elec, 3))).astype(np.float32)
    _, la = a.log_psi(th, r)
    logp = 2 * la
    times = []
    for i in range(warmup + steps):
        n = rng.standard_normal(r.shape).astype(np.float32)
        u = rng.random(n_walkers).astype(np.float32)
        t0 = time.perf_counter()
        r, logp, _, _, _ = a.walker_step(th, r, logp, n, u, SIGMA)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    v = n_walkers * len(times) / sum(times)
    return {"value": v, "unit": "walker-steps/s", "cores": cores, "kind": "port",
            "sample": f"{n_walkers} walkers x {len(times)} steps of the same H4 workload; torch.func restatement of the "
                      f"reference algorithm (nested-autodiff Hessian), float32, torch {torch.__version__}; "
                      "the reference's JAX path cannot run here (jax/flax/optax not installed)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    steps, warm = max(1, min(args.steps, 4)), max(1, min(args.warmup, 1))
    cb = cpu_baseline(args.cpu_walkers, steps=steps, warmup=warm)
    flops_step, _ = algorithmic_flops_per_walker_step(N_ATOMS, HIDDEN, N_ATOMS)
    print(json.dumps({
        "impl": "reference", "metric": "walker-steps/s", "value": cb["value"], "unit": "walker-steps/s",
        "n_gpus": world, "steps": steps, "warmup": warm, "ms_per_step": 1e3 * args.cpu_walkers / cb["value"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "H4 chain R=1.8 (2up/2dn) neural Slater-Jastrow, orbital MLP 11-64-64-1 x4, P=19976",
                   "walkers_per_step_sample": args.cpu_walkers, "algorithmic_flop_per_walker_step": flops_step},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "walker-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--walkers-per-gpu", type=int, default=WALKERS_PER_GPU)
    ap.add_argument("--burn", type=int, default=100)
    ap.add_argument("--cpu-walkers", type=int, default=4096)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip sustained / vmc_iteration / other_configs")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
