"""get_sigmas: host mirror of LiDARGen/models/__init__.py:5-18 (same expression, same dtype)."""
import numpy as np
import torch


def get_sigmas(config):
    if config.model.sigma_dist == 'geometric':
        sigmas = torch.tensor(
            np.exp(np.linspace(np.log(config.model.sigma_begin), np.log(config.model.sigma_end),
                               config.model.num_classes))).float().to(config.device)
    elif config.model.sigma_dist == 'uniform':
        sigmas = torch.tensor(
            np.linspace(config.model.sigma_begin, config.model.sigma_end, config.model.num_classes)
        ).float().to(config.device)
    else:
        raise NotImplementedError('sigma distribution not supported')
    return sigmas
