"""Sensor-model constants and sin/cos look-up tables of the cross-view step.

Host-side mirror of LiDARGen/models/KITTISampling.py:29-78,101-102 (identical block at
models/__init__.py:134-175): the same Python float64 expressions, so the constants handed to
the kernels are bit-identical to the reference's.  The LUTs are evaluated with torch on the
sample's device exactly like the reference's `torch.cos(azimuth)` (KITTISampling.py:176).
"""
import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SensorGeometry:
    H: int
    W: int
    R: int
    dh: float
    dv: float
    h_min: float
    v_min: float
    big_row_min: float
    cos_az: torch.Tensor
    sin_az: torch.Tensor
    cos_el: torch.Tensor
    sin_el: torch.Tensor


def sensor_geometry(H, W, device):
    h_scope_min, h_scope_max = -180, 180
    v_scope_max, v_scope_min = 3, -25        # "LIDARGEN's KITTI specs" (KITTISampling.py:46-47)
    h_scope = h_scope_max - h_scope_min
    v_scope = v_scope_max - v_scope_min
    dh = math.radians(h_scope) / W
    dv = math.radians(v_scope) / H
    h_min = ((W * h_scope_min) // h_scope) * dh + dh / 2
    R = int((np.max((np.absolute(v_scope_min), np.absolute(v_scope_max))) * 2) * H // v_scope)
    big_row_min = (R // -2) * dv + dv / 2
    v_min = ((H * v_scope_min) // v_scope) * dv + dv / 2
    az = torch.from_numpy(np.arange(W - 1, -1, -1) * dh + h_min).to(device)
    el = torch.from_numpy(np.arange(H - 1, -1, -1) * dv + v_min).to(device)
    return SensorGeometry(H, W, R, dh, dv, h_min, v_min, big_row_min,
                          torch.cos(az).contiguous(), torch.sin(az).contiguous(),
                          torch.cos(el).contiguous(), torch.sin(el).contiguous())
