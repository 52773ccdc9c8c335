"""Throughput of the SR kernels at production size: S = O^T O for P parameters x n samples (tcgen05 3xTF32 SYRK, TMA-fed),
then the CG solve.  Device time by CUDA events on the context's stream (torch only allocates here)."""
import os, sys, json
R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
import numpy as np
import torch
from _util import make_case, make_context

s, net, theta, _ = make_case(n_atoms=2, hidden=32, W=8, burn=0)
ctx = make_context(s, net)
ctx.set_params(theta)
dev = torch.device("cuda", ctx.device)
out = []
for P, n in [(4673, 65536), (19976, 65536), (19976, 16384)]:
    g = torch.Generator(device=dev).manual_seed(1)
    Ot = torch.randn((P, n), device=dev, dtype=torch.float32, generator=g)
    S = torch.empty((P, P), device=dev, dtype=torch.float32)
    torch.cuda.synchronize()
    st = torch.cuda.ExternalStream(ctx.stream, device=dev) if ctx.stream else torch.cuda.current_stream(dev)
    def run():
        rc = ctx.lib.nqmc_gram(ctx.h, Ot.data_ptr(), Ot.data_ptr(), P, P, n, n, n, 1.0 / n, 1e-3, 1, S.data_ptr(), P, ctx.stream)
        assert rc == 0, rc
    run(); ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(st):
        e0.record(st)
        for _ in range(3): run()
        e1.record(st)
    ctx.synchronize(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    alg = n * P * (P + 1.0)                      # upper triangle incl. diagonal, 2 flops per MAC
    tiles = sum(1 for tm in range((P + 127) // 128) for tn in range((P + 255) // 256) if tn * 256 + 256 > tm * 128)
    exe = tiles * 128 * 256 * 2.0 * n
    # spot check against torch fp64 on a corner
    ref = (Ot[:64].double() @ Ot[:64].double().T) / n + 1e-3 * torch.eye(64, device=dev, dtype=torch.float64)
    err = float((S[:64, :64].double() - ref).abs().max() / ref.abs().max())
    out.append({"P": P, "n": n, "ms": ms, "algorithmic_tflops": alg / ms / 1e9, "executed_tflops": exe / ms / 1e9,
                "operand_gb": P * n * 4 / 1e9, "max_rel_err_corner": err})
    print(json.dumps(out[-1]))
    del Ot, S
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(R, "gpurun_out", "sr_syrk_timing.json"), "w"), indent=1)
