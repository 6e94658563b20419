"""Noise schedule oracle (test infrastructure, see oracle/__init__.py).

Follows /root/reference/LiDARGen/models/__init__.py:5-18 (get_sigmas).
"""
import numpy as np
import torch


def sigma_schedule(sigma_begin, sigma_end, num_classes, dist="geometric"):
    """float32 tensor [num_classes]; float64 exp/linspace then one rounding to
    float32, exactly like the reference (models/__init__.py:7-9, 11-13)."""
    if dist == "geometric":
        s = np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), num_classes))
    elif dist == "uniform":
        s = np.linspace(sigma_begin, sigma_end, num_classes)
    else:
        raise NotImplementedError("sigma distribution not supported")
    return torch.tensor(s).float()
