"""One Stochastic Reconfiguration iteration for the benchmark wavefunction (H4, four 11-64-64-1 orbital networks,
P = 19,976) at 65,536 samples, stage by stage: E_L, per-walker O (nqmc_pergrad.cu), nqmc_sr_update (centre, SYRK, CG).
Device time by CUDA events on the context's stream; writes gpurun_out/sr_antisym_timing.json."""
import os, sys, json
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
import numpy as np
import torch
from _util import make_case, make_context

W = int(os.environ.get("NQ_W", 65536))
s, net, theta, r0 = make_case(n_atoms=4, hidden=64, W=2048, burn=20)
ctx = make_context(s, net)
ctx.set_params(theta)
rng = np.random.default_rng(1)
r = np.tile(r0, (W // 2048, 1, 1)) + 0.05 * rng.standard_normal((W, s.n_elec, 3)).astype(np.float32)
soa = ctx.to_soa(r.astype(np.float32))
dev = torch.device("cuda", ctx.device)
st = torch.cuda.ExternalStream(ctx.stream, device=dev) if ctx.stream else torch.cuda.current_stream(dev)

def timed(fn, reps=3):
    fn(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): out = fn()
    e1.record(st)
    ctx.synchronize(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

ms_e, e = timed(lambda: ctx.local_energy(soa))
ms_o, Ot = timed(lambda: ctx.log_psi_grads(soa))
P = ctx.P
res = {"P": int(P), "n": W, "ms_local_energy": ms_e, "ms_per_walker_O": ms_o, "O_gb": P * W * 4 / 1e9,
       "O_store_gbps": P * W * 4 / ms_o / 1e6}
def sr():
    O2 = ctx.log_psi_grads(soa)                      # sr_update centres O in place: a fresh O per repetition
    return ctx.sr_update(O2, e, 1e-3, 0.05, 100, 1e-6)
ms_all, (delta, info, S) = timed(sr, reps=2)
inf = info.numpy(ctx.stream)
res.update({"ms_O_plus_sr_update": ms_all, "ms_sr_update": ms_all - ms_o, "cg_iterations": float(inf[0]), "cg_residual": float(inf[1]),
            "delta_norm": float(np.linalg.norm(delta.numpy(ctx.stream)))})
print(json.dumps(res))
json.dump(res, open(os.path.join(R, "gpurun_out", "sr_antisym_timing.json"), "w"), indent=1)
