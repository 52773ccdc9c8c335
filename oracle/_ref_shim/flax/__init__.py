"""ORACLE / TEST INFRASTRUCTURE ONLY -- stand-in for the `flax` import name (see ../README.md)."""
from . import linen  # noqa: F401
