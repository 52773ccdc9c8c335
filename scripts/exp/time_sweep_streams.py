import os, sys
R="/root/repo"; sys.path[:0]=[R, R+"/tests", R+"/neural-qmc-wavefunction_b200"]
import numpy as np, torch
from nqmc_b200.sweep import GeometrySweep
for ns in (1, 2, 4, 8, 16):
    sw = GeometrySweep(bond_lengths=np.linspace(1.0, 4.0, 16), walkers_per_geometry=16384, hidden=(64, 64), step_size=0.5, seed=3, n_streams=ns)
    sw.burn_in(10)
    for _ in range(3): sw.step()
    sw.synchronize(); torch.cuda.synchronize()
    ev=[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a,b in ev:
        a.record(); sw.step(); b.record()
    torch.cuda.synchronize()
    ms=sum(a.elapsed_time(b) for a,b in ev)/5
    print(ns, "streams", ms, "ms/step", 16*16384/ms*1e3/1e6, "M walker-steps/s", [round(o["energy"],4) for o in sw.results()][:3])
    del sw
