from .antisymmetric import AntisymmetricWavefunction
from .base import BoundLogProb, Params, Wavefunction
from .simple import SimpleWavefunction

__all__ = ["Wavefunction", "Params", "SimpleWavefunction", "AntisymmetricWavefunction", "BoundLogProb"]
