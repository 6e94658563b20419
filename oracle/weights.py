"""Deterministic NCSN_LiDAR_small parameter sets (test infrastructure).

The inventory (names, shapes, registration order) restates the constructor at
/root/reference/LiDARGen/models/ncsnv2.py:420-477 with the block definitions of
models/layers.py:62-83 (CRP), 112-134 (RCU), 165-184 (MSF), 214-249 (Refine),
291-313 (ConvMeanPool), 401-456 (ResidualBlock) and
models/normalization.py:150-162 (InstanceNorm2dPlus: alpha, gamma, beta).

Values come from numpy's PCG64 seeded by crc32(name), so a fixture only has to
store inputs and outputs: the weights are regenerated bit-identically on any
box, independent of torch's RNG or module construction order.
"""
import zlib

import numpy as np
import torch

from .sigmas import sigma_schedule


def _residual_entries(prefix, cin, cout, kind):
    """kind: 'plain' (identity shortcut), 'down_pool' (res2.0), 'dilated' (res3.0/4.0)."""
    e = []
    if kind == "plain":
        e += [(f"{prefix}.conv1.weight", (cout, cin, 3, 3)), (f"{prefix}.conv1.bias", (cout,))]
        e += [(f"{prefix}.normalize2.{p}", (cout,)) for p in ("alpha", "gamma", "beta")]
        e += [(f"{prefix}.conv2.weight", (cout, cout, 3, 3)), (f"{prefix}.conv2.bias", (cout,))]
    elif kind == "down_pool":
        e += [(f"{prefix}.conv1.weight", (cin, cin, 3, 3)), (f"{prefix}.conv1.bias", (cin,))]
        e += [(f"{prefix}.normalize2.{p}", (cin,)) for p in ("alpha", "gamma", "beta")]
        e += [(f"{prefix}.conv2.conv.weight", (cout, cin, 3, 3)), (f"{prefix}.conv2.conv.bias", (cout,))]
        e += [(f"{prefix}.shortcut.conv.weight", (cout, cin, 1, 1)), (f"{prefix}.shortcut.conv.bias", (cout,))]
    elif kind == "dilated":
        e += [(f"{prefix}.conv1.weight", (cin, cin, 3, 3)), (f"{prefix}.conv1.bias", (cin,))]
        e += [(f"{prefix}.normalize2.{p}", (cin,)) for p in ("alpha", "gamma", "beta")]
        e += [(f"{prefix}.conv2.weight", (cout, cin, 3, 3)), (f"{prefix}.conv2.bias", (cout,))]
        e += [(f"{prefix}.shortcut.weight", (cout, cin, 3, 3)), (f"{prefix}.shortcut.bias", (cout,))]
    else:
        raise ValueError(kind)
    e += [(f"{prefix}.normalize1.{p}", (cin,)) for p in ("alpha", "gamma", "beta")]
    return e


def _refine_entries(prefix, in_planes, features, start=False, end=False):
    e = []
    for i, cp in enumerate(in_planes):
        for b in (1, 2):
            for s in (1, 2):
                e.append((f"{prefix}.adapt_convs.{i}.{b}_{s}_conv.weight", (cp, cp, 3, 3)))
    for b in range(1, (3 if end else 1) + 1):
        for s in (1, 2):
            e.append((f"{prefix}.output_convs.{b}_{s}_conv.weight", (features, features, 3, 3)))
    if not start:
        for i, cp in enumerate(in_planes):
            e.append((f"{prefix}.msf.convs.{i}.weight", (features, cp, 3, 3)))
            e.append((f"{prefix}.msf.convs.{i}.bias", (features,)))
    for i in (0, 1):
        e.append((f"{prefix}.crp.convs.{i}.weight", (features, features, 3, 3)))
    return e


def parameter_inventory(ngf=128, channels=2):
    """[(name, shape)] in the reference's registration order, parameters only
    (the 'sigmas' buffer is separate)."""
    g, g2 = ngf, 2 * ngf
    e = [("begin_conv.weight", (g, channels + 2, 3, 3)), ("begin_conv.bias", (g,))]
    e += [(f"normalizer.{p}", (g,)) for p in ("alpha", "gamma", "beta")]
    e += [("end_conv.weight", (channels, g, 3, 3)), ("end_conv.bias", (channels,))]
    e += _residual_entries("res1.0", g, g, "plain") + _residual_entries("res1.1", g, g, "plain")
    e += _residual_entries("res2.0", g, g2, "down_pool") + _residual_entries("res2.1", g2, g2, "plain")
    e += _residual_entries("res3.0", g2, g2, "dilated") + _residual_entries("res3.1", g2, g2, "plain")
    e += _residual_entries("res4.0", g2, g2, "dilated") + _residual_entries("res4.1", g2, g2, "plain")
    e += _refine_entries("refine1", [g2], g2, start=True)
    e += _refine_entries("refine2", [g2, g2], g2)
    e += _refine_entries("refine3", [g2, g2], g)
    e += _refine_entries("refine4", [g, g], g, end=True)
    return e


def make_state_dict(ngf=128, channels=2, num_classes=232, sigma_begin=50.0, sigma_end=0.01,
                    seed=1234, gain=1.0):
    """Deterministic float32 state_dict with the reference's key set (incl. 'sigmas').

    conv weights ~ U(-b, b), b = gain*sqrt(3/fan_in) (unit-variance preserving);
    biases ~ U(-0.1, 0.1); alpha/gamma ~ N(1, 0.02); beta ~ N(0, 0.02).
    """
    sd = {"sigmas": sigma_schedule(sigma_begin, sigma_end, num_classes)}
    for name, shape in parameter_inventory(ngf, channels):
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
        if name.endswith("weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            b = gain * np.sqrt(3.0 / fan_in)
            v = rng.uniform(-b, b, size=shape)
        elif name.endswith("bias"):
            v = rng.uniform(-0.1, 0.1, size=shape)
        elif name.endswith("alpha") or name.endswith("gamma"):
            v = rng.normal(1.0, 0.02, size=shape)
        else:
            v = rng.normal(0.0, 0.02, size=shape)
        sd[name] = torch.from_numpy(v.astype(np.float32))
    return sd
