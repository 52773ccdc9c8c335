"""Multi-GPU parity of the peer-memory communicator (csrc/nqmc_comm.cu): needs >= 2 GPUs on the box, skipped otherwise.
The single-GPU driver box skips it; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_peer_memory_allreduce_and_fused_step():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "scripts", "check_comm.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "CHECK_COMM PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_stochastic_reconfiguration_over_ranks():
    """Samples sharded over the ranks, S never reduced (q = S_r p is, every CG iteration): same update as the fp64 oracle
    on all samples, bitwise identical on every rank."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29534", os.path.join(ROOT, "scripts", "check_sr_multi.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert "CHECK_SR_MULTI PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_communicator_single_rank_is_identity_and_validates_arguments():
    """world = 1: create/attach succeed, the all-reduce is the identity; bad arguments are rejected, and the
    all-reduce refuses to run before a communicator is attached (no silent fallback)."""
    import torch
    from nqmc_b200 import _native
    from _util import make_case, make_context
    s, net, theta, r = make_case(n_atoms=2, hidden=64, W=256, burn=2)
    ctx = make_context(s, net)
    ctx.set_params(torch.from_numpy(theta).cuda())
    v = torch.arange(8, dtype=torch.float64, device="cuda")
    with pytest.raises(_native.NativeError):
        ctx.comm_allreduce(v.clone())                      # nothing attached yet
    with pytest.raises(_native.NativeError):
        ctx.comm_create(3, 2)                              # rank out of range
    h = ctx.comm_create(0, 1)
    assert len(h) == ctx.COMM_HANDLE_BYTES
    ctx.comm_attach([h])
    w = v.clone()
    ctx.comm_allreduce(w)
    torch.cuda.synchronize()
    assert torch.equal(w, v)
    with pytest.raises(_native.NativeError):
        ctx.comm_allreduce(torch.zeros(ctx.P + 64, dtype=torch.float64, device="cuda"))   # longer than the inbox
