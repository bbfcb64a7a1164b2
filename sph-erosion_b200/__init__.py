"""sph-erosion_b200: B200-native (sm_100a) implementation of SPH-Erosion's per-step particle hot path
behind the reference's FluidSystemSPH / Grid surface.  Import with
importlib.import_module("sph-erosion_b200") (the directory name carries the upstream repo name)."""
from . import build, capi  # noqa: F401
from .fluid import FluidSystemSPH  # noqa: F401
from .grid import Grid  # noqa: F401

__all__ = ["FluidSystemSPH", "Grid", "build", "capi"]
