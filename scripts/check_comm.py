"""Multi-GPU check of the peer-memory communicator (csrc/nqmc_comm.cu).  Run under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_comm.py
1. nqmc_comm_allreduce vs NCCL all_reduce on random fp64 vectors (tolerance: summation order) and bitwise equality
   of the result across ranks;  2. fused nqmc_walker_step (in-stream all-reduces) vs the rank-local halves + NCCL;
3. 300 back-to-back exchanges of alternating sizes (flag / parity reuse)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "neural-qmc-wavefunction_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from nqmc_b200 import parallel  # noqa: E402
import bench  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
from _util import T  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
mol, wf, params, theta = bench.make_problem()
W = 4096
lo = rank * W


def make_ctx():
    c = wf.context(mol)
    c.set_params(theta)
    c.reserve(W)
    return c


ctx = make_ctx()
assert parallel.attach_peer_memory(ctx)
ok = True
# 1 ------------------------------------------------------------------------------------------------
g = torch.Generator(device="cpu").manual_seed(100 + rank)
for n in (8, 1000, ctx.P):
    v = torch.randn(n, dtype=torch.float64, generator=g).to(dev)
    ref = v.clone()
    dist.all_reduce(ref)
    ctx.comm_allreduce(v)
    torch.cuda.synchronize()
    err = float((v - ref).abs().max() / ref.abs().max())
    allv = [torch.empty_like(v) for _ in range(world)]
    dist.all_gather(allv, v)
    same = all(torch.equal(allv[0], a) for a in allv)
    if rank == 0:
        print(f"allreduce n={n}: rel err vs NCCL {err:.2e}, bitwise equal across ranks: {same}")
    ok &= err < 1e-13 and same
# 3 ------------------------------------------------------------------------------------------------
v8 = torch.full((8,), float(rank + 1), dtype=torch.float64, device=dev)
vp = torch.full((ctx.P,), float(rank + 1), dtype=torch.float64, device=dev)
expect = world * (world + 1) / 2
for it in range(300):
    a, b = v8.clone(), vp.clone()
    ctx.comm_allreduce(a)
    ctx.comm_allreduce(b)
    if it % 50 == 49:
        torch.cuda.synchronize()
        ok &= bool((a == expect).all()) and bool((b == expect).all())
torch.cuda.synchronize()
if rank == 0:
    print("300 back-to-back exchanges:", "ok" if ok else "MISMATCH")
# timing: back-to-back exchanges, peer memory vs NCCL (device time, both include rank skew)
for name, fn in (("peer", ctx.comm_allreduce), ("nccl", dist.all_reduce)):
    for vec in (v8.clone(), vp.clone()):
        for _ in range(20):
            fn(vec)
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(200):
            fn(vec)
        e1.record()
        torch.cuda.synchronize()
        if rank == 0:
            print(f"{name} all-reduce of {vec.numel()} f64: {e0.elapsed_time(e1) / 200 * 1e3:.2f} us per call")
# 2 ------------------------------------------------------------------------------------------------
ref_ctx = make_ctx()
outs = []
for c, fused in ((ctx, True), (ref_ctx, False)):
    r = c.to_soa(bench.initial_walkers(mol, W, lo))
    _, la = c.log_psi(r)
    logp = (2.0 * T(la)).contiguous()
    e_l = torch.empty(W, dtype=torch.float32, device=dev)
    hdr = torch.empty(8, dtype=torch.float64, device=dev)
    gsum = torch.empty(c.P, dtype=torch.float64, device=dev)
    for s in range(3):
        if fused:
            c.walker_step(r, logp, 0.3, e_l, hdr, gsum, seed=7, step=s, walker_id0=lo)
        else:
            c.walker_step_a(r, logp, 0.3, e_l, hdr, seed=7, step=s, walker_id0=lo)
            dist.all_reduce(hdr)
            c.walker_step_b(r, e_l, hdr, gsum)
            dist.all_reduce(gsum)
    torch.cuda.synchronize()
    outs.append((hdr.cpu().numpy().copy(), gsum.cpu().numpy().copy(), np.array(r.cpu().numpy())))
(h1, g1, r1), (h2, g2, r2) = outs
eh = np.abs(h1 - h2).max() / np.abs(h2).max()
eg = np.abs(g1 - g2).max() / np.abs(g2).max()
if rank == 0:
    print(f"fused step vs halves+NCCL: hdr rel {eh:.2e}  gsum rel {eg:.2e}  walkers equal {np.array_equal(r1, r2)}  n={h1[0]:.0f}")
ok &= eh < 1e-12 and eg < 1e-6 and np.array_equal(r1, r2) and h1[0] == world * W
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("CHECK_COMM", "PASS" if flag.item() == 1.0 else "FAIL")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
