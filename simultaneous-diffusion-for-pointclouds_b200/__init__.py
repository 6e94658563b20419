"""B200-native simultaneous multi-view Langevin sampling step.

Host side: Python mirror of the reference's sampler / score-network interface
(LiDARGen/models) on top of the C ABI in include/sdpc_b200.h.  Importable as `sdpc_b200`
through the shim module at the repository root (the directory name is not an identifier).
"""
__version__ = "0.1.0"
