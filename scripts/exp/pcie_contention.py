"""Host<->device copy bandwidth of the e2e step's buffers (3.4 MB H2D, 3.8 MB D2H per rank, pinned memory), one rank at
a time versus all ranks at once (torchrun, one process per GPU).  Names the limiter of the host-buffer path at 8 GPUs."""
import json, os, time
import torch, torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H2D, D2H = 3407872, 3829888
hi, ho = torch.empty(H2D, dtype=torch.uint8).pin_memory(), torch.empty(D2H, dtype=torch.uint8).pin_memory()
di, do = torch.empty(H2D, dtype=torch.uint8, device="cuda"), torch.empty(D2H, dtype=torch.uint8, device="cuda")
s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def h2d():
    di.copy_(hi, non_blocking=True)


def d2h():
    ho.copy_(do, non_blocking=True)


def both():
    with torch.cuda.stream(s_in):
        di.copy_(hi, non_blocking=True)
    with torch.cuda.stream(s_out):
        ho.copy_(do, non_blocking=True)


res = {}
for name, fn in (("h2d", h2d), ("d2h", d2h), ("both", both)):
    alone = None
    for turn in range(world):                                    # one rank at a time
        barrier()
        if turn == rank:
            alone = timed(fn)
    barrier()
    together = timed(fn)                                         # all ranks at once
    barrier()
    t = torch.tensor([alone, together], dtype=torch.float64, device="cuda")
    if world > 1:
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
    else:
        tl = [t]
    res[name] = {"alone_us": [round(float(x[0]), 1) for x in tl], "together_us": [round(float(x[1]), 1) for x in tl]}
if rank == 0:
    res["bytes"] = {"h2d": H2D, "d2h": D2H}
    res["n_gpus"] = world
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
