"""The three annealed-Langevin samplers of the reference, re-hosted on the C ABI.

Signatures (positional order, defaults, return values) are the reference's:
  a-4 anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti  LiDARGen/models/KITTISampling.py:6-513
  a-5 anneal_Langevin_dynamics_inpainting_simultaneous_basic        LiDARGen/models/__init__.py:112-602
  a-6 anneal_Langevin_dynamics_inpainting                           LiDARGen/models/__init__.py:1385-1442
The level / step loops, the RNG draw (`torch.randn_like`, so the Philox stream is the
reference's) and the list-of-CPU-tensor returns stay in Python; each Langevin update and the
whole cross-view block is one `sdpc_langevin_reproject_step` call.  CUDA tensors only.
"""
import numpy as np
import torch

from . import cabi
from .step import StepRunner, translation_origins


def _snap(t):
    """CPU snapshot like the reference's `.to('cpu')`, but never an alias of a live buffer."""
    return t.detach().to('cpu', copy=True)


def _require_cuda(x):
    if not x.is_cuda:
        raise cabi.SdpcError("the B200 samplers need CUDA tensors: there is no CPU fallback")


def _report(prefix_verbose, c, step_size, sigma, grad, grad_likelihood, noise, x_mod, refer_image, grad_ref):
    """The reference's per-level diagnostics (KITTISampling.py:147-149,492-500)."""
    grad_norm = torch.norm(grad.view(grad.shape[0], -1), dim=-1).mean()
    grad_likelihood_norm = torch.norm(grad_likelihood.view(grad.shape[0], -1), dim=-1).mean()
    noise_norm = torch.norm(noise.view(noise.shape[0], -1), dim=-1).mean()
    image_norm = torch.norm(x_mod.view(x_mod.shape[0], -1), dim=-1).mean()
    snr = np.sqrt(step_size / 2.) * grad_norm / noise_norm
    grad_mean_norm = torch.norm(grad.mean(dim=0).view(-1)) ** 2 * sigma ** 2
    print("grad_ref: {}, mean: {}, median: {}".format(grad_ref, torch.mean(torch.abs(x_mod - refer_image)),
                                                      torch.median(torch.abs(x_mod - refer_image))))
    print("level: {}, step_size: {}, grad_norm: {}, grad_likelihood_norm: {}, image_norm: {}, snr: {}, grad_mean_norm: {}".format(
        c, step_size, grad_norm.item(), grad_likelihood_norm.item(), image_norm.item(), snr.item(), grad_mean_norm.item()))


def _simultaneous(variant, x_mod, refer_image, refer_mask, sky, minStepToShare, setting, allowance, scorenet, sigmas,
                  actualBatchSize, n_steps_each, step_lr, existMask, denoise, verbose, grad_ref,
                  correlation_coefficient, fromWorld=None, toWorld=None, modificationList=None, shard=None,
                  _lib=None):
    if _lib is None:            # _lib: test-only host emulation of the C ABI (tests/host_emul), CPU tensors
        _require_cuda(x_mod)
    dev = x_mod.device
    images, targets, sharedImages = [], [], []
    B = x_mod.shape[0]
    x = x_mod.detach().to(torch.float32).clone().contiguous()       # the reference rebinds x_mod out of place
    refer = refer_image.to(device=dev, dtype=torch.float32)
    kw = {}
    if variant == cabi.SDPC_VARIANT_POSE:
        kw = dict(to_world=toWorld.to(dev).reshape(B, 4, 4), from_world=fromWorld.to(dev).reshape(B, 4, 4))
    else:
        kw = dict(origins=translation_origins(modificationList.to(dev)))
    run = StepRunner(x.shape, dev, refer, refer_mask, sky, existMask, actualBatchSize, variant, lib=_lib, **kw)
    if shard is not None:
        shard.attach(run, x)
    grad_full = torch.zeros_like(x) if shard is not None else None
    pose = variant == cabi.SDPC_VARIANT_POSE
    L = len(sigmas)
    new_images = torch.empty_like(x)
    grad_likelihood = torch.zeros_like(x)
    mask_f = None
    for c, sigma in enumerate(sigmas):
        if pose:                                               # KITTISampling.py:107-110
            if setting == 6:
                correlation_coefficient = 1 / (L / (c + 1))
            if setting == 7:
                correlation_coefficient = 0.5 / (L / (c + 1))
        else:                                                  # models/__init__.py:210-213
            if setting == 5:
                correlation_coefficient = 1 / (L / (c + 1))
            if setting == 6:
                correlation_coefficient = 0.5 / (L / (c + 1))
        sigmaMod = 1
        if sigma > 1:
            sigmaMod = sigma
        labels = torch.ones(B, device=dev) * c
        labels = labels.long()
        step_size = step_lr * (sigma / sigmas[-1]) ** 2         # numpy-scalar arithmetic, as the reference
        noise_scale = np.sqrt(step_size * 2)
        share = c >= minStepToShare
        if pose:
            p = run.params(step_size, noise_scale, grad_ref, correlation_coefficient, sigmaMod, share,
                           min_depth_filter=(setting == 5), allowance=allowance, sky_filter=False)
        else:
            allow = (5 if setting >= 8 else 10) if setting >= 7 else None
            p = run.params(step_size, noise_scale, grad_ref, correlation_coefficient, sigmaMod, share,
                           min_depth_filter=True, allowance=allow, sky_filter=True)
        want_images = share and (c in (0, 20, 110) or c == L - 1)
        wants_print = (verbose and c % 20 == 0) or (pose and (c == 1 or c == 2))
        for s in range(n_steps_each):
            grad = scorenet(x, labels) if shard is None else shard.score(scorenet, x, labels, grad_full)
            noise = torch.randn_like(x)        # sharded: every rank draws the full tensor (same seed) and uses its block
            last = s == n_steps_each - 1
            keep_gl = last and (c == L - 1 or wants_print)
            b = run.buffers(x, grad, noise, grad_likelihood=grad_likelihood if keep_gl else None,
                            new_images=new_images if want_images else None)
            if shard is None:
                run.step(p, b)
            else:
                shard.step(run, p, b, x)
            if want_images:
                snap = _snap(new_images if shard is None else shard.gather_result(new_images))
                if c in (0, 20, 110):
                    sharedImages.append(snap)
                if c == L - 1:
                    images.append(snap)
        if wants_print and shard is None:
            _report(verbose, c, step_size, sigma, torch.nan_to_num(grad), grad_likelihood, noise, x, refer, grad_ref)

    if mask_f is None:
        mask_f = refer_mask.to(dev)
    if shard is not None:                                       # finish on the rank's own views, then gather
        sl = slice(shard.lo, shard.hi)
        xl, rl, ml = x[sl], refer[sl], mask_f[sl]
        if denoise:
            last_noise = ((L - 1) * torch.ones(xl.shape[0], device=dev)).long()
            xl = xl + sigmas[-1] ** 2 * scorenet(xl.contiguous(), last_noise) + grad_ref * grad_likelihood[sl]
        xl = xl + grad_ref * (-ml * (xl - rl))
        x[sl] = xl
        images.append(_snap(shard.gather_result(x)))
        return images, targets, sharedImages
    if denoise:                                                 # KITTISampling.py:502-507 (stale grad_likelihood)
        last_noise = ((L - 1) * torch.ones(B, device=dev)).long()
        x = x + sigmas[-1] ** 2 * scorenet(x, last_noise) + grad_ref * grad_likelihood
        if pose:
            print("grad_ref: {}, mean: {}, median: {}".format(grad_ref, torch.mean(torch.abs(x - refer)),
                                                              torch.median(torch.abs(x - refer))))
        else:
            print("grad_ref: {}, mean: {}, median: {}".format(grad_ref, torch.mean(torch.abs(x - refer)),
                                                              torch.median(torch.abs(x - refer))))
    grad_likelihood = -mask_f * (x - refer)
    x = x + grad_ref * grad_likelihood
    images.append(_snap(x))
    return images, targets, sharedImages


@torch.no_grad()
def anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
        x_mod, refer_image, refer_mask, sky, x_indices, minStepToShare, setting, allowance, scorenet, sigmas,
        fromWorld, toWorld, actualBatchSize, n_steps_each=100, step_lr=0.000008, existMask=None, denoise=True,
        verbose=True, grad_ref=0.1, correlation_coefficient=0.1, sampling_step=16, shard=None, _lib=None):
    """a-4, pose-matrix simultaneous sampler (Line.yml path).  `shard`: optional dist.ViewShard."""
    return _simultaneous(cabi.SDPC_VARIANT_POSE, x_mod, refer_image, refer_mask, sky, minStepToShare, setting,
                         allowance, scorenet, sigmas, actualBatchSize, n_steps_each, step_lr, existMask, denoise,
                         verbose, grad_ref, correlation_coefficient, fromWorld=fromWorld, toWorld=toWorld, shard=shard,
                         _lib=_lib)


@torch.no_grad()
def anneal_Langevin_dynamics_inpainting_simultaneous_basic(
        x_mod, refer_image, refer_mask, sky, x_indices, minStepToShare, setting, scorenet, sigmas, modificationList,
        actualBatchSize, n_steps_each=100, step_lr=0.000008, existMask=None, denoise=True, verbose=True,
        grad_ref=0.1, correlation_coefficient=0.1, sampling_step=16, shard=None, _lib=None):
    """a-5, translation-only simultaneous sampler (Inpainting / Densification path)."""
    return _simultaneous(cabi.SDPC_VARIANT_TRANSLATION, x_mod, refer_image, refer_mask, sky, minStepToShare, setting,
                         None, scorenet, sigmas, actualBatchSize, n_steps_each, step_lr, existMask, denoise, verbose,
                         grad_ref, correlation_coefficient, modificationList=modificationList, shard=shard, _lib=_lib)


@torch.no_grad()
def anneal_Langevin_dynamics_inpainting(x_mod, refer_image, refer_mask, scorenet, sigmas, n_steps_each=100,
                                        step_lr=0.000008, denoise=True, verbose=True, grad_ref=0.1, sampling_step=16):
    """a-6, single-view baseline: Langevin update only, a CPU snapshot after every step
    (models/__init__.py:1422) and no nan_to_num on the score."""
    _require_cuda(x_mod)
    dev = x_mod.device
    images, targets = [], []
    B = x_mod.shape[0]
    x = x_mod.detach().to(torch.float32).clone().contiguous()
    refer = refer_image.to(device=dev, dtype=torch.float32)
    run = StepRunner(x.shape, dev, refer, refer_mask, None, None, 1, cabi.SDPC_VARIANT_POSE)
    L = len(sigmas)
    grad_likelihood = torch.zeros_like(x)
    for c, sigma in enumerate(sigmas):
        labels = (torch.ones(B, device=dev) * c).long()
        step_size = step_lr * (sigma / sigmas[-1]) ** 2
        noise_scale = np.sqrt(step_size * 2)
        p = run.params(step_size, noise_scale, grad_ref, 0.0, 1.0, False, False, None, False, nan_to_num=False)
        wants_print = verbose and c % 20 == 0
        for s in range(n_steps_each):
            grad = scorenet(x, labels)
            noise = torch.randn_like(x)
            last = s == n_steps_each - 1
            keep_gl = last and (c == L - 1 or wants_print)
            b = run.buffers(x, grad, noise, grad_likelihood=grad_likelihood if keep_gl else None)
            run.update_only(p, b)
            images.append(_snap(x))
        if wants_print:
            _report(verbose, c, step_size, sigma, grad, grad_likelihood, noise, x, refer, grad_ref)
    mask_d = refer_mask.to(dev)
    if denoise:
        last_noise = ((L - 1) * torch.ones(B, device=dev)).long()
        x = x + sigmas[-1] ** 2 * scorenet(x, last_noise) + grad_ref * grad_likelihood
        images.append(_snap(x))
        print("grad_ref: {}, mean: {}, median: {}".format(grad_ref, torch.mean(torch.abs(x - refer)),
                                                          torch.median(torch.abs(x - refer))))
    grad_likelihood = -mask_d * (x - refer)
    x = x + grad_ref * grad_likelihood
    images.append(_snap(x))
    print("grad_ref: {}, mean: {}, median: {}".format(grad_ref, torch.mean(torch.abs(x - refer)),
                                                      torch.median(torch.abs(x - refer))))
    targets.append(_snap(refer))
    return images, targets


# ---- row N4: the unconditional and the beam-densification samplers, on the same update kernel ----------
@torch.no_grad()
def anneal_Langevin_dynamics(x_mod, scorenet, sigmas, n_steps_each=200, step_lr=0.000008,
                             final_only=False, verbose=False, denoise=True):
    """Unconditional annealed Langevin sampling (LiDARGen/models/__init__.py:20-58): x += eps*s + sqrt(2 eps)*z.
    Runs `sdpc_langevin_update` with an all-zero mask and grad_ref = 0 (the likelihood term is then exactly +-0)."""
    _require_cuda(x_mod)
    dev = x_mod.device
    images = []
    B = x_mod.shape[0]
    x = x_mod.detach().to(torch.float32).clone().contiguous()
    run = StepRunner(x.shape, dev, torch.zeros_like(x), torch.zeros(x.shape, dtype=torch.int32, device=dev), None, None, 1,
                     cabi.SDPC_VARIANT_POSE)
    L = len(sigmas)
    for c, sigma in enumerate(sigmas):
        labels = (torch.ones(B, device=dev) * c).long()
        step_size = step_lr * (sigma / sigmas[-1]) ** 2
        p = run.params(step_size, np.sqrt(step_size * 2), 0.0, 0.0, 1.0, False, False, None, False, nan_to_num=False)
        for s in range(n_steps_each):
            grad = scorenet(x, labels)
            noise = torch.randn_like(x)
            run.update_only(p, run.buffers(x, grad, noise))
            if not final_only:
                images.append(_snap(x))
            if verbose:
                flat = lambda t: t.view(t.shape[0], -1)
                grad_norm, noise_norm = torch.norm(flat(grad), dim=-1).mean(), torch.norm(flat(noise), dim=-1).mean()
                print("level: {}, step_size: {}, grad_norm: {}, image_norm: {}, snr: {}, grad_mean_norm: {}".format(
                    c, step_size, grad_norm.item(), torch.norm(flat(x), dim=-1).mean().item(),
                    (np.sqrt(step_size / 2.) * grad_norm / noise_norm).item(),
                    (torch.norm(grad.mean(dim=0).view(-1)) ** 2 * sigma ** 2).item()))
    if denoise:
        last_noise = ((L - 1) * torch.ones(B, device=dev)).long()
        x = x + sigmas[-1] ** 2 * scorenet(x, last_noise)
        images.append(_snap(x))
    return [_snap(x)] if final_only else images


@torch.no_grad()
def anneal_Langevin_dynamics_densification(x_mod, refer_image, scorenet, sigmas, n_steps_each=100, step_lr=0.000008,
                                           denoise=True, verbose=True, grad_ref=0.1, sampling_step=16):
    """Beam densification of one view (LiDARGen/models/__init__.py:60-109): beams 0, sampling_step, ... are known, the
    rest is sampled.  Same loop as a-6 with the mask built here (the reference's unused bilinear `raw_interp`, :66, is
    not evaluated)."""
    mask = torch.zeros(x_mod.shape, dtype=torch.int32, device=x_mod.device)
    mask[:, :, 0:64:sampling_step, :] = 1
    return anneal_Langevin_dynamics_inpainting(x_mod, refer_image, mask, scorenet, sigmas, n_steps_each, step_lr, denoise,
                                               verbose, grad_ref)
