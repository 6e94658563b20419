"""Noise schedule of the samplers: host mirror of `get_sigmas` (LiDARGen/models/__init__.py:5-18).

The values are the bit-exactness contract of row a-7 (tests/golden/sigmas.npz, recorded from the unmodified reference):
levels spaced evenly in float64 - in log(sigma) for the 'geometric' schedule, in sigma for 'uniform' - and rounded to
float32 once at the end."""
import numpy as np
import torch

_SPACING = {
    # name -> (map onto the axis the levels are evenly spaced on, map back)
    "geometric": (np.log, np.exp),
    "uniform": (lambda v: v, lambda v: v),
}


def noise_levels(sigma_begin, sigma_end, num_classes, dist="geometric"):
    """float64 numpy array [num_classes], largest noise level first."""
    if dist not in _SPACING:
        raise NotImplementedError('sigma distribution not supported')
    forward, back = _SPACING[dist]
    return back(np.linspace(forward(sigma_begin), forward(sigma_end), num_classes))


def get_sigmas(config):
    """float32 tensor [config.model.num_classes] on config.device (the reference's call signature)."""
    m = config.model
    levels = noise_levels(m.sigma_begin, m.sigma_end, m.num_classes, m.sigma_dist)
    return torch.tensor(levels).float().to(config.device)
