"""saras_b200 -- B200-native multigrid pressure-Poisson path of SARAS (lib/poisson + its halo exchange).

The product is the CUDA library behind include/sarasmg.h plus the C++ drop-in for the reference's
`poisson` / `multigrid_d2` / `multigrid_d3` classes under saras_b200/host/.  This Python package is
plumbing for tests and bench.py (ctypes over the same C-ABI).
"""
from . import capi, host_api  # noqa: F401

__all__ = ["capi", "host_api"]
