"""nqmc_b200 -- B200-native walker loop behind the nqmc call signatures.

Sub-packages mirror the reference's layout (src/nqmc/{systems,wavefunction,sampler,hamiltonian,
optimizer}); ``import nqmc_b200 as nqmc`` is the intended drop-in for the VMC hot path.
Importing this package does not touch CUDA; the first wavefunction evaluation loads
libnqmc_b200.so and raises if it (or a B200) is missing -- there is no CPU fallback.
"""
from . import analysis, checkpoint, hamiltonian, optimizer, parallel, params, random, sampler, systems, wavefunction
from ._native import Context, NativeError, load_library

__version__ = "0.1.0"
__all__ = ["analysis", "systems", "wavefunction", "sampler", "hamiltonian", "optimizer", "parallel", "params", "random", "checkpoint",
           "Context", "NativeError", "load_library"]
