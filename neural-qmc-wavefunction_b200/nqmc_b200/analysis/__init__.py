"""Statistical analysis of VMC samples (honest error bars)."""
from .statistics import blocking_analysis, chain_blocking, compute_autocorrelation, effective_sample_size, reblocking_curve

__all__ = ["blocking_analysis", "chain_blocking", "compute_autocorrelation", "effective_sample_size", "reblocking_curve"]
