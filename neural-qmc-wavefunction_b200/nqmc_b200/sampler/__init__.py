from .metropolis import MetropolisSampler, SamplerState

__all__ = ["MetropolisSampler", "SamplerState"]
