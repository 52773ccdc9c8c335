from .molecular import (EPSILON, electron_electron_potential, electron_nuclear_potential, kinetic_energy, local_energy,
                        local_energy_batched)

__all__ = ["kinetic_energy", "electron_nuclear_potential", "electron_electron_potential", "local_energy",
           "local_energy_batched", "EPSILON"]
