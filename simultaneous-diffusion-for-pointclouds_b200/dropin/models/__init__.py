"""Drop-in replacement for the reference's `models` package on the sampling hot path.

Put this directory's parent (`.../dropin`) ahead of `LiDARGen/` on sys.path and the reference
runners' own imports resolve to the B200 implementation unchanged
(runners/ncsn_runner_kitti_simultaneous.py:15-27):

    from models import (anneal_Langevin_dynamics_inpainting,
                        anneal_Langevin_dynamics_inpainting_simultaneous_basic, get_sigmas)
    from models.KITTISampling import anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti
    from models.ncsnv2 import NCSN_LiDAR_small
    from models.ema import EMAHelper
"""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)
import sdpc_b200  # noqa: E402,F401
from sdpc_b200.samplers import (anneal_Langevin_dynamics, anneal_Langevin_dynamics_densification,  # noqa: E402,F401
                                anneal_Langevin_dynamics_inpainting,
                                anneal_Langevin_dynamics_inpainting_simultaneous_basic)
from sdpc_b200.sigmas import get_sigmas  # noqa: E402,F401
