"""Exponential-moving-average weights for the score network.

Interface and checkpoint layout of the reference's helper (LiDARGen/models/ema.py:4-46): `shadow` maps
un-prefixed parameter names to tensors and is what a checkpoint stores last (`states[-1]`,
runners/ncsn_runner_kitti_simultaneous.py:485-489).  Sampling only needs `register` / `load_state_dict` /
`ema`; `update` and `ema_copy` are kept so that training-side callers keep working.
"""
import torch
import torch.nn as nn


def _trainable(module):
    """(name, parameter) pairs of the wrapped network, DataParallel unwrapped, frozen parameters skipped."""
    inner = module.module if isinstance(module, nn.DataParallel) else module
    return inner, [(n, p) for n, p in inner.named_parameters() if p.requires_grad]


class EMAHelper(object):
    def __init__(self, mu=0.999):
        self.mu = mu
        self.shadow = {}

    def register(self, module):
        _, params = _trainable(module)
        self.shadow = {name: p.detach().clone() for name, p in params}

    @torch.no_grad()
    def update(self, module):
        _, params = _trainable(module)
        keep = self.mu
        for name, p in params:
            # (1 - mu) * theta + mu * shadow with one rounding per product and one for the sum, like the reference
            # expression (models/ema.py:20); torch.add(..., alpha=mu) would fuse the second product into the sum
            fresh, old = p.detach() * (1. - keep), self.shadow[name] * keep
            self.shadow[name] = fresh + old

    @torch.no_grad()
    def ema(self, module):
        """overwrite the live parameters with the averaged ones (what sampling does after loading a checkpoint)."""
        inner, params = _trainable(module)
        for name, p in params:
            p.copy_(self.shadow[name].to(p.device))
        refresh = getattr(inner, "refresh_weights", None)
        if refresh is not None:
            refresh()                      # the CUDA handle keeps repacked copies of the weights

    def ema_copy(self, module):
        inner, _ = _trainable(module)
        twin = type(inner)(inner.config).to(inner.config.device)
        twin.load_state_dict(inner.state_dict())
        if isinstance(module, nn.DataParallel):
            twin = nn.DataParallel(twin)
        self.ema(twin)
        return twin

    def state_dict(self):
        return self.shadow

    def load_state_dict(self, state_dict):
        self.shadow = state_dict
