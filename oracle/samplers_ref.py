"""Annealed-Langevin sampler oracles (test infrastructure, see oracle/__init__.py).

Loops restated from
  a-4  /root/reference/LiDARGen/models/KITTISampling.py:6-513   (pose matrices)
  a-5  /root/reference/LiDARGen/models/__init__.py:112-602      (translations)
  a-6  /root/reference/LiDARGen/models/__init__.py:1385-1442    (single view)
The per-step math lives in crossview_ref.py.  `noise_fn(x)` replaces
torch.randn_like so that tests can inject identical noise into both sides.
"""
import numpy as np
import torch

from . import crossview_ref as cv


def _step_constants(step_lr, sigma, sigma_last):
    """step_size and noise scale with the reference's numpy-scalar arithmetic
    (KITTISampling.py:135,156): float32 under NumPy>=2, see SURVEY quirk (xi)."""
    step_size = step_lr * (sigma / sigma_last) ** 2
    return step_size, np.sqrt(step_size * 2)


def langevin_update(x, grad, refer, mask, noise, step_size, noise_scale, grad_ref):
    """x + eps*grad + rho*(-mask*(x-ref)) + sqrt(2 eps)*z, left to right, fp32 (KITTISampling.py:144-156)."""
    grad_likelihood = -mask * (x - refer)
    return x + step_size * grad + grad_ref * grad_likelihood + noise * noise_scale, grad_likelihood


@torch.no_grad()
def sampler_pose(x_mod, refer_image, refer_mask, sky, x_indices, minStepToShare, setting, allowance,
                 scorenet, sigmas, fromWorld, toWorld, actualBatchSize, n_steps_each=100,
                 step_lr=0.000008, existMask=None, denoise=True, verbose=True, grad_ref=0.1,
                 correlation_coefficient=0.1, noise_fn=torch.randn_like, trace=None):
    """a-4.  Same positional signature as the reference up to `correlation_coefficient`."""
    images, targets, shared = [], [], []
    dev = x_mod.device
    sky = sky.to(dev)
    fromWorld = torch.squeeze(fromWorld).to(dev)
    toWorld = torch.squeeze(toWorld).to(dev)
    geo = cv.make_geometry(x_mod.shape[-2], x_mod.shape[-1], dev)
    L = len(sigmas)
    grad_likelihood = None
    for c, sigma in enumerate(sigmas):
        if setting == 6:
            correlation_coefficient = 1 / (L / (c + 1))
        if setting == 7:
            correlation_coefficient = 0.5 / (L / (c + 1))
        sigma_mod = sigma if sigma > 1 else 1
        labels = (torch.ones(x_mod.shape[0], device=dev) * c).long()
        step_size, noise_scale = _step_constants(step_lr, sigma, sigmas[-1])
        for s in range(n_steps_each):
            grad = torch.nan_to_num(scorenet(x_mod, labels))
            noise = noise_fn(x_mod)
            x_mod, grad_likelihood = langevin_update(x_mod, grad, refer_image, refer_mask, noise,
                                                     step_size, noise_scale, grad_ref)
            if c >= minStepToShare:
                new_images, image_mask, too_high = cv.shared_images(
                    x_mod, geo, sigma_mod, actualBatchSize, existMask, sky,
                    to_world=toWorld, from_world=fromWorld,
                    min_depth_filter=(setting == 5), controlled_average=True, allowance=allowance)
                if c in (0, 20, 110):
                    shared.append(new_images.to("cpu"))
                if c == L - 1:
                    images.append(new_images.to("cpu"))
                x_mod = cv.apply_correction(x_mod, new_images, image_mask, sky, refer_mask, too_high,
                                            correlation_coefficient)
                if trace is not None:
                    trace.append(dict(c=c, s=s, new_images=new_images.cpu(), x=x_mod.cpu()))
    if denoise:
        last = ((L - 1) * torch.ones(x_mod.shape[0], device=dev)).long()
        x_mod = x_mod + sigmas[-1] ** 2 * scorenet(x_mod, last) + grad_ref * grad_likelihood
    x_mod = x_mod + grad_ref * (-refer_mask * (x_mod - refer_image))
    images.append(x_mod.to("cpu"))
    return images, targets, shared


@torch.no_grad()
def sampler_translation(x_mod, refer_image, refer_mask, sky, x_indices, minStepToShare, setting,
                        scorenet, sigmas, modificationList, actualBatchSize, n_steps_each=100,
                        step_lr=0.000008, existMask=None, denoise=True, verbose=True, grad_ref=0.1,
                        correlation_coefficient=0.1, noise_fn=torch.randn_like, trace=None):
    """a-5."""
    images, targets, shared = [], [], []
    dev = x_mod.device
    sky = sky.to(dev)
    geo = cv.make_geometry(x_mod.shape[-2], x_mod.shape[-1], dev)
    origins = cv.translation_origins(modificationList.to(dev))
    L = len(sigmas)
    grad_likelihood = None
    for c, sigma in enumerate(sigmas):
        if setting == 5:
            correlation_coefficient = 1 / (L / (c + 1))
        if setting == 6:
            correlation_coefficient = 0.5 / (L / (c + 1))
        sigma_mod = sigma if sigma > 1 else 1
        labels = (torch.ones(x_mod.shape[0], device=dev) * c).long()
        step_size, noise_scale = _step_constants(step_lr, sigma, sigmas[-1])
        for s in range(n_steps_each):
            grad = torch.nan_to_num(scorenet(x_mod, labels))
            noise = noise_fn(x_mod)
            x_mod, grad_likelihood = langevin_update(x_mod, grad, refer_image, refer_mask, noise,
                                                     step_size, noise_scale, grad_ref)
            if c >= minStepToShare:
                new_images, image_mask, too_high = cv.shared_images(
                    x_mod, geo, sigma_mod, actualBatchSize, existMask, sky, origins=origins,
                    min_depth_filter=True, controlled_average=(setting >= 7),
                    allowance=(5.0 if setting >= 8 else 10.0), sky_filter=True)
                if c in (0, 20, 110):
                    shared.append(new_images.to("cpu"))
                if c == L - 1:
                    images.append(new_images.to("cpu"))
                x_mod = cv.apply_correction(x_mod, new_images, image_mask, sky, refer_mask, too_high,
                                            correlation_coefficient)
                if trace is not None:
                    trace.append(dict(c=c, s=s, new_images=new_images.cpu(), x=x_mod.cpu()))
    if denoise:
        last = ((L - 1) * torch.ones(x_mod.shape[0], device=dev)).long()
        x_mod = x_mod + sigmas[-1] ** 2 * scorenet(x_mod, last) + grad_ref * grad_likelihood
    x_mod = x_mod + grad_ref * (-refer_mask * (x_mod - refer_image))
    images.append(x_mod.to("cpu"))
    return images, targets, shared


@torch.no_grad()
def sampler_single_view(x_mod, refer_image, refer_mask, scorenet, sigmas, n_steps_each=100,
                        step_lr=0.000008, denoise=True, verbose=True, grad_ref=0.1,
                        noise_fn=torch.randn_like):
    """a-6: no nan_to_num, a CPU snapshot after every step."""
    images, targets = [], []
    dev = x_mod.device
    L = len(sigmas)
    grad_likelihood = None
    for c, sigma in enumerate(sigmas):
        labels = (torch.ones(x_mod.shape[0], device=dev) * c).long()
        step_size, noise_scale = _step_constants(step_lr, sigma, sigmas[-1])
        for s in range(n_steps_each):
            grad = scorenet(x_mod, labels)
            noise = noise_fn(x_mod)
            x_mod, grad_likelihood = langevin_update(x_mod, grad, refer_image, refer_mask, noise,
                                                     step_size, noise_scale, grad_ref)
            images.append(x_mod.to("cpu"))
    if denoise:
        last = ((L - 1) * torch.ones(x_mod.shape[0], device=dev)).long()
        x_mod = x_mod + sigmas[-1] ** 2 * scorenet(x_mod, last) + grad_ref * grad_likelihood
        images.append(x_mod.to("cpu"))
    x_mod = x_mod + grad_ref * (-refer_mask * (x_mod - refer_image))
    images.append(x_mod.to("cpu"))
    targets.append(refer_image.to("cpu"))
    return images, targets


@torch.no_grad()
def sampler_unconditional(x_mod, scorenet, sigmas, n_steps_each=200, step_lr=0.000008, final_only=False,
                          denoise=True, noise_fn=torch.randn_like):
    """row N4, anneal_Langevin_dynamics (LiDARGen/models/__init__.py:20-58): x += eps*s + sqrt(2 eps)*z, no likelihood
    term, no nan_to_num; a CPU snapshot per step unless final_only."""
    images = []
    dev = x_mod.device
    L = len(sigmas)
    for c, sigma in enumerate(sigmas):
        labels = (torch.ones(x_mod.shape[0], device=dev) * c).long()
        step_size, noise_scale = _step_constants(step_lr, sigma, sigmas[-1])
        for s in range(n_steps_each):
            grad = scorenet(x_mod, labels)
            noise = noise_fn(x_mod)
            x_mod = x_mod + step_size * grad + noise * noise_scale          # :35
            if not final_only:
                images.append(x_mod.to("cpu"))
    if denoise:
        last = ((L - 1) * torch.ones(x_mod.shape[0], device=dev)).long()
        x_mod = x_mod + sigmas[-1] ** 2 * scorenet(x_mod, last)             # :51
        images.append(x_mod.to("cpu"))
    return [x_mod.to("cpu")] if final_only else images


@torch.no_grad()
def sampler_densification(x_mod, refer_image, scorenet, sigmas, n_steps_each=100, step_lr=0.000008, denoise=True,
                          grad_ref=0.1, sampling_step=16, noise_fn=torch.randn_like):
    """row N4, anneal_Langevin_dynamics_densification (LiDARGen/models/__init__.py:60-109): the known pixels are the
    beams 0, sampling_step, 2*sampling_step, ... (mask built inside, float); otherwise the a-6 loop including the stale
    likelihood gradient in the denoise step (:96)."""
    mask = torch.zeros_like(x_mod)
    mask[:, :, 0:64:sampling_step, :] = 1                                   # :67 (the bilinear `raw_interp` of :66 is unused)
    return sampler_single_view(x_mod, refer_image, mask, scorenet, sigmas, n_steps_each, step_lr, denoise, False,
                               grad_ref, noise_fn)
