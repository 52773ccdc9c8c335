"""Multi-GPU check of Stochastic Reconfiguration (nqmc_sr_update with the peer-memory communicator attached).  Run under
torchrun:  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/check_sr_multi.py
The samples are sharded over the ranks; every rank must end with the same update as the fp64 oracle computes from ALL
samples (oracle/sr.py, the restatement of .github/phase6-issue-body.md:150-181), bitwise identical across ranks."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

R = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path[:0] = [R, os.path.join(R, "tests"), os.path.join(R, "neural-qmc-wavefunction_b200")]
from nqmc_b200 import parallel  # noqa: E402
from oracle.sr import sr_update  # noqa: E402
from oracle.layout import KIND_SIMPLE  # noqa: E402
from _util import make_case, make_context  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ["NQMC_DEVICE"] = str(local)
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
s, net, theta, r = make_case(n_atoms=2, hidden=32, kind=KIND_SIMPLE, W=8, burn=0)
ctx = make_context(s, net)
ctx.set_params(theta)
ctx.reserve(8)
assert parallel.attach_peer_memory(ctx)
P = ctx.P
ok = True
for n_local, reg in ((512, 1e-3), (1000, 1e-2)):
    n = n_local * world
    rng = np.random.default_rng(11)                              # the same full problem on every rank
    mix = np.eye(P) + 0.3 * rng.standard_normal((P, P)) / np.sqrt(P)
    g = (rng.standard_normal((n, P)) @ mix + 2.0 * rng.standard_normal(P)).astype(np.float32)
    e = (-1.1 + 0.3 * rng.standard_normal(n) + 0.05 * g[:, 0]).astype(np.float32)
    e[5] = np.nan                                                 # one node hit, on rank 0's shard
    lo = rank * n_local
    Ot = ctx.upload(np.ascontiguousarray(g[lo:lo + n_local].T))
    delta, info, _ = ctx.sr_update(Ot, ctx.upload(e[lo:lo + n_local]), reg=reg, lr=0.05, max_iter=400, tol=1e-9)
    d = torch.as_tensor(delta.numpy(ctx.stream)).cuda()
    keep = np.isfinite(e)
    ref = sr_update(g[keep], e[keep], learning_rate=0.05, regularization=reg)
    err = float(np.linalg.norm(d.cpu().numpy() - ref) / np.linalg.norm(ref))
    alld = [torch.empty_like(d) for _ in range(world)]
    dist.all_gather(alld, d)
    same = all(torch.equal(alld[0], a) for a in alld)
    inf = info.numpy(ctx.stream)
    if rank == 0:
        print(f"SR over {world} ranks, P={P}, n={n}: rel err vs fp64 oracle {err:.2e}, CG iterations {inf[0]:.0f}, "
              f"residual {inf[1]:.1e}, bitwise equal across ranks: {same}")
    ok &= err < 2e-4 and same and inf[2] == 1.0
t = torch.tensor([1.0 if ok else 0.0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("CHECK_SR_MULTI PASS" if t.item() == 1.0 else "CHECK_SR_MULTI FAIL")
dist.barrier()
dist.destroy_process_group()
