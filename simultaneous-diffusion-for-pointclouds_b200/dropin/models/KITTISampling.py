"""models.KITTISampling shim (reference: LiDARGen/models/KITTISampling.py:6)."""
from sdpc_b200.samplers import anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti  # noqa: F401
