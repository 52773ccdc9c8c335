from .sr import StochasticReconfiguration
from .vmc import VMCOptimizer, VMCState, compute_gradient, estimate_energy

__all__ = ["VMCOptimizer", "VMCState", "compute_gradient", "estimate_energy", "StochasticReconfiguration"]
