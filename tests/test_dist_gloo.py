"""World-size-2 gloo test of the multi-GPU host logic: walkers sharded by global id, the two additive
exchanges (hdr, gsum) all-reduced, gradient combined -- must equal the single-process estimator.
Per-shard sums come from the CPU oracle here (the device kernels need a GPU); the collective plumbing,
sharding and combination are the code the GPU path uses (nqmc_b200.parallel)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, W, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, "neural-qmc-wavefunction_b200")]
    from nqmc_b200 import parallel
    from oracle.layout import OracleNet, hydrogen_chain_system, init_theta
    from oracle.nqmc_oracle import Oracle
    from oracle import philox
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = hydrogen_chain_system(4)
    net = OracleNet(hidden=(8, 8))
    th = init_theta(s, net, 1)
    o = Oracle(s, net)
    lo, hi = parallel.shard_range(W, rank, world)
    ids = np.arange(lo, hi)
    # positions and streams are functions of the GLOBAL walker id => independent of the rank count
    rng_r = np.random.default_rng(0).standard_normal((W, 4, 3))[lo:hi]
    r = s.R[np.arange(4) % 4][None] + 0.5 * rng_r
    normals, uniforms, _ = philox.walker_streams(7, 3, ids, 4)
    logp = 2 * o.log_psi(th, r)[1]
    st = o.walker_step(th, r, logp, normals.astype(np.float64), uniforms.astype(np.float64), 0.3)
    e = st["e_l"]
    hdr = torch.tensor([len(e), e.sum(), (e * e).sum(), st["accept"].sum(), 0, 0, 0, 0], dtype=torch.float64)
    parallel.all_reduce_sum_(hdr)                                   # exchange 1
    mean = float(hdr[1] / hdr[0])
    gsum = torch.tensor(((e - mean)[:, None] * st["O"]).sum(0))
    parallel.all_reduce_sum_(gsum)                                  # exchange 2
    g, m, sd = parallel.combine_gradient(hdr.numpy(), gsum.numpy())
    assert parallel.world() == (rank, world)
    if rank == 0:
        np.savez(out, g=g, mean=m, std=sd, nacc=float(hdr[3]))
    dist.destroy_process_group()


@pytest.mark.parametrize("W", [64, 37])
def test_two_rank_gradient_equals_single_process(tmp_path, W):
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), W, out), nprocs=2, join=True)
    got = np.load(out)
    from oracle.layout import OracleNet, hydrogen_chain_system, init_theta
    from oracle.nqmc_oracle import Oracle
    from oracle import philox
    s = hydrogen_chain_system(4)
    net = OracleNet(hidden=(8, 8))
    th = init_theta(s, net, 1)
    o = Oracle(s, net)
    r = s.R[np.arange(4) % 4][None] + 0.5 * np.random.default_rng(0).standard_normal((W, 4, 3))
    normals, uniforms, _ = philox.walker_streams(7, 3, np.arange(W), 4)
    st = o.walker_step(th, r, 2 * o.log_psi(th, r)[1], normals.astype(np.float64), uniforms.astype(np.float64), 0.3)
    e, O = st["e_l"], st["O"]
    g_ref = 2 * np.mean((e - e.mean())[:, None] * O, axis=0)      # vmc.py:126-133
    assert np.allclose(got["g"], g_ref, rtol=1e-10, atol=1e-13)
    assert np.isclose(got["mean"], e.mean()) and np.isclose(got["std"], e.std() / np.sqrt(W))
    assert got["nacc"] == st["accept"].sum()
