"""Sampling runners: thin re-host of the callers of the hot path.

Mirrors `sample()` of LiDARGen/runners/ncsn_runner_kitti_simultaneous.py:461-924 (Line.yml, pose-matrix
sampler) and LiDARGen/runners/ncsn_runner_AllForOne.py:466-1000 (Inpainting.yml / Densification.yml,
translation sampler): checkpoint + EMA loading, existTotal mask preprocessing, the `doThis` ablation loop
with the reference's hard-coded hyper-parameters, output post-processing and .npy naming.  It is a boundary
row of SURVEY.md 8 (keep the API), not a kernel target; the data source is the synthetic generator unless
`b200.data_root` points at KITTI-360 (`datasets.KITTI360Line` / `KITTI360AllForOne` / `KITTI360Densification`, row N2 of
SURVEY 8f).
"""
import logging
import os
import time

import numpy as np
import torch

from .ema import EMAHelper
from .samplers import (anneal_Langevin_dynamics_inpainting, anneal_Langevin_dynamics_inpainting_simultaneous_basic,
                       anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti)
from .scorenet import NCSN_LiDAR_small
from .sigmas import get_sigmas
from .synthetic_data import SyntheticMultiView

__all__ = ["NCSNRunnerKITTISimultaneous", "NCSNRunnerAllForOne", "get_model"]


def get_model(config):
    """every LiDAR dataset maps to NCSN_LiDAR_small (ncsn_runner_kitti_simultaneous.py:33-52)."""
    prec = getattr(getattr(config, "b200", None), "precision", None)
    return NCSN_LiDAR_small(config, precision=prec).to(config.device)


def inverse_data_transform(config, X):
    """datasets/__init__.py:203-215 for logit_transform=False, rescaled=False: clamp to [0, 1]."""
    return torch.clamp(X, 0.0, 1.0)


def _erode(mask, iterations):
    """scipy.ndimage.binary_erosion(border_value=1) with the default cross structuring element."""
    try:
        import scipy.ndimage
        return scipy.ndimage.binary_erosion(mask, border_value=1, iterations=iterations)
    except Exception:                                       # pragma: no cover
        m = mask.copy()
        for _ in range(iterations):
            p = np.pad(m, 1, constant_values=True)
            m = p[1:-1, 1:-1] & p[:-2, 1:-1] & p[2:, 1:-1] & p[1:-1, :-2] & p[1:-1, 2:]
        return m


def exist_mask(config, batch_size):
    """ncsn_runner_kitti_simultaneous.py:527-533: threshold at max/3, 4x erosion of rows 2.., tile."""
    path = getattr(getattr(config, "b200", None), "exist_mask", None) or "/data/existTotalLiDARGenSettings.npy"
    H, W = config.data.image_size, config.data.image_width
    if os.path.exists(path) and np.load(path).shape == (H, W):
        vals = np.load(path)
    else:                                                   # synthetic: sensor drop-outs near the image border rows
        r = np.random.Generator(np.random.PCG64(7))
        vals = r.uniform(0.4, 1.0, size=(H, W)) * 8601.0
        vals[: max(1, H // 16)] *= 0.2
    vals = vals > np.max(vals) / 3
    vals[2:] = _erode(vals[2:], 4)
    vals = np.tile(np.expand_dims(vals, 0), (batch_size, 1, 1))
    return torch.from_numpy(vals).to(config.device)


def _to_grid_layout(x):
    """[B,2,H,W] -> channel-major [2B,3,H,W] as the runners save it (ncsn_runner_kitti_simultaneous.py:848-870)."""
    x = x.transpose(1, 0)
    x = x.reshape((x.size(1) * x.size(0), 1, x.size(2), x.size(3)))
    return torch.cat((x, x, x), 1)


class _Base:
    def __init__(self, args, config):
        self.args, self.config = args, config
        args.log_sample_path = os.path.join(args.log_path, "samples") if hasattr(args, "log_path") else None

    # ---- model -----------------------------------------------------------------------------------
    def load_score(self):
        cfg = self.config
        score = torch.nn.DataParallel(get_model(cfg), device_ids=[torch.device(cfg.device).index or 0])
        ckpt = getattr(getattr(cfg, "b200", None), "checkpoint", None)
        if ckpt:
            states = torch.load(ckpt, map_location=cfg.device)
            score.load_state_dict(states[0], strict=True)
            if cfg.model.ema:
                ema_helper = EMAHelper(mu=cfg.model.ema_rate)
                ema_helper.register(score)
                ema_helper.load_state_dict(states[-1])
                ema_helper.ema(score)
        else:
            logging.info("no b200.checkpoint configured: random-init weights (offline run)")
        score.eval()
        return score

    def dataset(self, mode):
        cfg = self.config
        root = getattr(getattr(cfg, "b200", None), "data_root", None)
        if root:                                            # KITTI-360 on disk: row N2's GPU dataset assembly
            from . import datasets
            reader = {"line": datasets.KITTI360Line, "allforone": datasets.KITTI360AllForOne,
                      "densification": datasets.KITTI360Densification}[mode]
            return datasets.ItemBatches(reader(root, cfg, device=cfg.device), cfg.sampling.batch_size)
        return SyntheticMultiView(cfg.data.image_size, cfg.data.image_width, cfg.sampling.batch_size,
                                  cfg.sampling.actualBatchSize, mode=mode, seed=self.args.seed)

    def save_grid(self, images, nrow, name):
        """PNG rendering of a [N,3,H,W] stack, `make_grid` + `save_image` like the reference (presentation only: a missing
        torchvision is logged, the .npy arrays next to it are the data)."""
        try:
            from torchvision.utils import make_grid, save_image
            save_image(make_grid(images.cpu().detach().float(), max(1, int(nrow))), os.path.join(self.args.image_folder, name))
        except ImportError as e:                        # torchvision or its PIL back end
            logging.warning("PNG grid %s not written: %s", name, e)

    def save_inputs(self, doThis, saveNum, grid_tag, refer_images_full, refer_mask_full, goalImages, refer_sky):
        """Known pixels, ground truth and sky mask of the batch, written once per batch (doThis == 0):
        ncsn_runner_kitti_simultaneous.py:650-696 (grids named by batch number), ncsn_runner_AllForOne.py:662-711 (by ids)."""
        cfg, out = self.config, self.args.image_folder
        tag = '_{}.pth'.format(cfg.sampling.ckpt_id)
        nrow = int(np.sqrt(cfg.sampling.batch_size))
        known = _to_grid_layout(inverse_data_transform(cfg, refer_images_full * refer_mask_full))
        self.save_grid(known, nrow, str(doThis) + '_' + grid_tag + '_Input_image_grid_{}.png'.format(cfg.sampling.ckpt_id))
        np.save(os.path.join(out, str(doThis) + '_' + saveNum + '_Input_completion' + tag), known.cpu().detach().numpy())
        goal = _to_grid_layout(inverse_data_transform(cfg, goalImages))
        self.save_grid(goal, nrow, str(doThis) + '_' + grid_tag + '_GT_image_grid_{}.png'.format(cfg.sampling.ckpt_id))
        np.save(os.path.join(out, str(doThis) + '_' + saveNum + '_GT_completion' + tag), goal.cpu().detach().numpy())
        np.save(os.path.join(out, str(doThis) + '_' + saveNum + '_SKY' + tag), refer_sky.clone().cpu().detach().numpy())

    def save_outputs(self, doThis, saveNum, n_views, all_outputs, grid_tag=None, nrow=None, shared_initial=False):
        """`all_outputs[-1]` as `<doThis>_<ids>_Masked_completion_<ckpt>.pth.npy` (+ PNG grid); with `shared_initial` also
        `all_outputs[-2]` as `..._Shared_completion_initial<ckpt>.pth.npy` (ncsn_runner_AllForOne.py:911,976-994)."""
        cfg = self.config
        shp = (n_views, cfg.data.channels, cfg.data.image_size, cfg.data.image_width)
        sample = inverse_data_transform(cfg, all_outputs[-1].view(*shp))
        masked = _to_grid_layout(sample)
        np.save(os.path.join(self.args.image_folder,
                             str(doThis) + '_' + saveNum + '_Masked_completion_{}.pth'.format(cfg.sampling.ckpt_id)),
                masked.cpu().detach().numpy())
        nrow = int(np.sqrt(cfg.sampling.batch_size)) if nrow is None else nrow
        grid_tag = saveNum if grid_tag is None else grid_tag
        self.save_grid(masked, nrow, str(doThis) + '_' + grid_tag + '_Masked_image_grid_{}.png'.format(cfg.sampling.ckpt_id))
        if shared_initial:
            first = _to_grid_layout(inverse_data_transform(cfg, all_outputs[-2].view(*shp)))
            np.save(os.path.join(self.args.image_folder,
                                 str(doThis) + '_' + saveNum + '_Shared_completion_initial{}.pth'.format(cfg.sampling.ckpt_id)),
                    first.cpu().detach().numpy())
            self.save_grid(first, nrow, str(doThis) + '_' + saveNum + '_Shared_image_grid_initial{}.png'.format(cfg.sampling.ckpt_id))
        return masked


class NCSNRunnerKITTISimultaneous(_Base):
    """Line.yml: `KITTI360_im_8batch`, pose-matrix sampler (main.py:191-192)."""

    def sample(self):
        cfg, args = self.config, self.args
        score = self.load_score()
        sigmas = get_sigmas(cfg).cpu().numpy()
        A, Bsz = cfg.sampling.actualBatchSize, cfg.sampling.batch_size
        data = self.dataset("line")
        existVals = exist_mask(cfg, Bsz)
        n_batches = int(getattr(getattr(cfg, "b200", None), "max_batches", 1) or 1)
        timeTaken = np.zeros(max(A, n_batches))
        for batchesToDo in range(n_batches):
            (refer_images_full, refer_mask_full, refer_sky, refer_indices_full, toWorld_full, fromWorld_full, goalImages,
             toOGView, saveNumArray) = data.batch(batchesToDo)
            G = Bsz // A
            saveNum = "".join(str(saveNumArray[p * A].cpu().detach().numpy()) + "_" for p in range(G))
            np.save(os.path.join(args.image_folder, "toWorld_" + saveNum), toWorld_full.cpu().detach())
            np.save(os.path.join(args.image_folder, "fromWorld_" + saveNum), toOGView.cpu().detach())
            for doThis in range(A):
                startStep, correlation_co, gradRef, allowance, setting = 2, 0.01, 1, 10, 5      # :574-579
                dev = cfg.device
                refer = refer_images_full.float().to(dev)
                refer_mask = refer_mask_full.int().to(dev)
                sky, idx = refer_sky.clone(), refer_indices_full.clone()
                toWorld, fromWorld = toWorld_full.clone(), fromWorld_full.clone()
                init_samples = torch.rand(Bsz, cfg.data.channels, cfg.data.image_size, cfg.data.image_width, device=dev)
                if doThis == 0:
                    self.save_inputs(doThis, saveNum, str(batchesToDo), refer_images_full, refer_mask_full, goalImages,
                                     refer_sky)
                start_time = time.time()
                if doThis == A - 1:                                                           # LiDARGen baseline arm
                    n_views = Bsz
                    all_outputs, _ = anneal_Langevin_dynamics_inpainting(
                        init_samples, refer, refer_mask, score, sigmas, cfg.sampling.n_steps_each, cfg.sampling.step_lr,
                        denoise=cfg.sampling.denoise, grad_ref=1, sampling_step=4)
                else:
                    k = doThis + 2 if doThis < A - 2 else A                                   # views kept per group

                    def keep(t, shape):
                        t = torch.reshape(t, (G, A, -1))
                        return torch.reshape(t[:, :k], (G * k,) + shape)
                    img = (cfg.data.channels, cfg.data.image_size, cfg.data.image_width)
                    one = (1, cfg.data.image_size, cfg.data.image_width)
                    n_views = G * k
                    all_outputs, _, _ = anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
                        keep(init_samples, img), keep(refer, img), keep(refer_mask, img), keep(sky, one), keep(idx, one),
                        startStep, setting, allowance, score, sigmas, keep(fromWorld, (4, 4)), keep(toWorld, (4, 4)), k,
                        cfg.sampling.n_steps_each, cfg.sampling.step_lr, existMask=existVals,
                        denoise=cfg.sampling.denoise, grad_ref=gradRef, correlation_coefficient=correlation_co,
                        sampling_step=4)
                torch.cuda.synchronize()
                timeTaken[doThis] += (time.time() - start_time)
                print("--- %s seconds ---" % (timeTaken[doThis] / (batchesToDo + 1)))
                np.save(os.path.join(args.image_folder, str(doThis) + '_' + saveNum + '_TimeTaken.npy'), timeTaken[doThis])
                self.save_outputs(doThis, saveNum, n_views, all_outputs, grid_tag=str(batchesToDo),
                                  nrow=int(np.sqrt((doThis + 2) * G)) if doThis < A - 2 else None)       # :872-884
        return 0


class NCSNRunnerAllForOne(_Base):
    """Inpainting.yml / Densification.yml: translation sampler (ncsn_runner_AllForOne.py:466-1000)."""

    def sample(self):
        cfg, args = self.config, self.args
        score = self.load_score()
        sigmas = get_sigmas(cfg).cpu().numpy()
        A, Bsz = cfg.sampling.actualBatchSize, cfg.sampling.batch_size
        dens = cfg.data.dataset == 'KITTI360_im_simultaneous_densification'
        data = self.dataset("densification" if dens else "allforone")
        existVals = exist_mask(cfg, Bsz)
        n_batches = int(getattr(getattr(cfg, "b200", None), "max_batches", 1) or 1)
        timeTaken = np.zeros(max(A, n_batches))
        for batchesToDo in range(n_batches):
            (refer_images_full, refer_mask_full, refer_sky, refer_indices_full, toWorld_full, fromWorld_full, goalImages,
             toOGView, saveNumArray) = data.batch(batchesToDo)
            G = Bsz // A
            saveNum = "".join(str(saveNumArray[p * A].cpu().detach().numpy()) + "_" for p in range(G))
            np.save(os.path.join(args.image_folder, "toWorld_" + saveNum), toWorld_full.cpu().detach())
            np.save(os.path.join(args.image_folder, "fromWorld_" + saveNum), toOGView.cpu().detach())
            endPoint, toAdd = (2, A - 2) if dens else (A, 0)                                   # :555-558
            for doThis in range(endPoint):
                startStep, correlation_co, gradRef, setting = 2, 0.01, 1, 7                    # :585-594
                dev = cfg.device
                modifiersToGive = torch.from_numpy(np.array(cfg.data.modifications)).to(dev)
                refer = refer_images_full.float().to(dev)
                refer_mask = refer_mask_full.int().to(dev)
                sky, idx = refer_sky.clone(), refer_indices_full.clone()
                init_samples = torch.rand(Bsz, cfg.data.channels, cfg.data.image_size, cfg.data.image_width, device=dev)
                img = (cfg.data.channels, cfg.data.image_size, cfg.data.image_width)
                one = (1, cfg.data.image_size, cfg.data.image_width)
                if doThis == 0:
                    self.save_inputs(doThis, saveNum, saveNum, refer_images_full, refer_mask_full, goalImages, refer_sky)

                def keep(t, shape, k):
                    t = torch.reshape(t, (G, A, -1))
                    return torch.reshape(t[:, :k], (G * k,) + shape)
                start_time = time.time()
                if doThis + toAdd == A - 1:                                                    # baseline on view 0 of each group
                    n_views = G
                    all_outputs, _ = anneal_Langevin_dynamics_inpainting(
                        keep(init_samples, img, 1), keep(refer, img, 1), keep(refer_mask, img, 1), score, sigmas,
                        cfg.sampling.n_steps_each, cfg.sampling.step_lr, denoise=cfg.sampling.denoise, grad_ref=1,
                        sampling_step=4)
                else:
                    k = doThis + 2 if doThis + toAdd < A - 2 else A
                    n_views = G * k
                    all_outputs, _, _ = anneal_Langevin_dynamics_inpainting_simultaneous_basic(
                        keep(init_samples, img, k), keep(refer, img, k), keep(refer_mask, img, k), keep(sky, one, k),
                        keep(idx, one, k), startStep, setting, score, sigmas, modifiersToGive, k,
                        cfg.sampling.n_steps_each, cfg.sampling.step_lr, existMask=existVals,
                        denoise=cfg.sampling.denoise, grad_ref=gradRef, correlation_coefficient=correlation_co,
                        sampling_step=4)
                torch.cuda.synchronize()
                timeTaken[doThis] += (time.time() - start_time)
                print("--- %s seconds ---" % (timeTaken[doThis] / (batchesToDo + 1)))
                np.save(os.path.join(args.image_folder, str(doThis) + '_' + saveNum + '_TimeTaken.npy'), timeTaken[doThis])
                if doThis + toAdd == A - 1:                                                    # :944-994
                    nrow = int(np.sqrt(G))
                elif doThis + toAdd < A - 2:
                    nrow = int(np.sqrt((doThis + 2) * G))
                else:
                    nrow = None
                self.save_outputs(doThis, saveNum, n_views, all_outputs, nrow=nrow, shared_initial=True)
        return 0
