"""EMAHelper with the reference's interface and checkpoint layout (LiDARGen/models/ema.py:4-46):
`shadow` maps un-prefixed parameter names to tensors (states[-1] of a checkpoint)."""
import torch.nn as nn


def _unwrap(module):
    return module.module if isinstance(module, nn.DataParallel) else module


class EMAHelper(object):
    def __init__(self, mu=0.999):
        self.mu = mu
        self.shadow = {}

    def register(self, module):
        for name, param in _unwrap(module).named_parameters():
            if param.requires_grad:
                self.shadow[name] = param.data.clone()

    def update(self, module):
        for name, param in _unwrap(module).named_parameters():
            if param.requires_grad:
                self.shadow[name].data = (1. - self.mu) * param.data + self.mu * self.shadow[name].data

    def ema(self, module):
        inner = _unwrap(module)
        for name, param in inner.named_parameters():
            if param.requires_grad:
                param.data.copy_(self.shadow[name].data)
        if hasattr(inner, "refresh_weights"):
            inner.refresh_weights()          # the CUDA handle keeps packed copies of the weights

    def ema_copy(self, module):
        inner = _unwrap(module)
        module_copy = type(inner)(inner.config).to(inner.config.device)
        module_copy.load_state_dict(inner.state_dict())
        if isinstance(module, nn.DataParallel):
            module_copy = nn.DataParallel(module_copy)
        self.ema(module_copy)
        return module_copy

    def state_dict(self):
        return self.shadow

    def load_state_dict(self, state_dict):
        self.shadow = state_dict
