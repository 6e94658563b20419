"""Marshalling of one Langevin + cross-view step onto the C ABI (include/sdpc_b200.h).

`StepRunner` owns the static per-call state of a sampler invocation (masks, poses, LUTs,
z-buffer workspace) and issues `sdpc_langevin_reproject_step` on the current CUDA stream.
"""
import ctypes as C

import numpy as np
import torch

from . import cabi
from .geometry import sensor_geometry


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def min_depth_threshold(sigma_mod):
    """float32 value of torch.log2(torch.tensor(0.2)+1)/6*sigmaMod (KITTISampling.py:273-274)."""
    return float(torch.log2(torch.tensor(0.2) + 1) / 6 * sigma_mod)


def translation_origins(modification_list):
    """a-5's originList (models/__init__.py:224-231), same fp32 torch ops as the reference."""
    og = torch.unsqueeze(torch.unsqueeze(modification_list, -1), -1)
    o = (torch.log2(torch.abs(og) + 1)) / 6
    o = torch.pow(2, (o * 6)) - 1
    return o / (og + 0.00000001) * 10


class StepRunner:
    def __init__(self, x_shape, device, refer, mask, sky, exist, group_size, variant,
                 to_world=None, from_world=None, origins=None, lib=None, tgt_first=0, tgt_count=0,
                 debug=False, scalar_div_recip=None):
        B, Cn, H, W = x_shape
        assert Cn == 2
        if B % group_size != 0:
            raise ValueError(f"batch of {B} views is not a multiple of actualBatchSize={group_size}")
        if exist is not None and exist.shape[0] < group_size:
            raise ValueError(f"existMask has {exist.shape[0]} views, actualBatchSize={group_size} are indexed")
        self.lib = lib if lib is not None else cabi.load()
        self.device = device
        self.B, self.A, self.H, self.W = B, group_size, H, W
        self.geo = sensor_geometry(H, W, device)
        self.variant = variant
        f32 = dict(device=device, dtype=torch.float32)
        # The kernels index refer / mask as dense [B,2,H,W] buffers (float4 / int4 loads) and sky / exist as dense byte
        # planes: broadcastable inputs (the reference's `-mask*(x-refer)` accepts e.g. a [B,1,H,W] mask) are expanded
        # here, anything that cannot broadcast to the sample's shape is rejected before a device pointer is formed.
        self.refer = self._expand(refer, x_shape, "refer_image").to(**f32).contiguous()
        self.mask = self._expand(mask, x_shape, "refer_mask").to(device=device, dtype=torch.int32).contiguous()
        if sky is not None and sky.numel() != B * H * W:
            raise ValueError(f"sky has {sky.numel()} elements, expected B*H*W = {B * H * W} ([B,1,H,W])")
        if exist is not None and tuple(exist.shape[1:]) != (H, W):
            raise ValueError(f"existMask is {tuple(exist.shape)}, expected [>={group_size},{H},{W}]")
        self.sky = sky.to(device=device).reshape(B, H, W).to(torch.uint8).contiguous() if sky is not None else None
        self.exist = exist[:group_size].to(device=device).to(torch.uint8).contiguous() if exist is not None else None
        for name, t in (("toWorld", to_world), ("fromWorld", from_world)):
            if t is not None and t.numel() != B * 16:
                raise ValueError(f"{name} has {t.numel()} elements, expected B*16 = {B * 16} ([B,1,4,4])")
        if origins is not None and (origins.dim() != 4 or origins.shape[0] < group_size or origins.shape[1] != 3):
            raise ValueError(f"originList is {tuple(origins.shape)}, expected [>={group_size},3,1,1]")
        self.to_world = to_world.to(device=device, dtype=torch.float64).reshape(B, 16).contiguous() if to_world is not None else None
        self.from_world = from_world.to(device=device, dtype=torch.float64).reshape(B, 16).contiguous() if from_world is not None else None
        self.origins = origins[:group_size, :, 0, 0].to(**f32).contiguous() if origins is not None else None
        self.tgt_first, self.tgt_count = tgt_first, tgt_count
        # torch's CUDA kernels evaluate tensor / python-scalar as tensor * (1/scalar); its CPU kernels divide.
        # Follow whichever reference arm runs on this device (override for tests against CPU goldens).
        self.scalar_div_recip = (torch.device(device).type == "cuda") if scalar_div_recip is None else bool(scalar_div_recip)
        nbytes = 256 if self.lib is None else self._ws_bytes()
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=device)
        init = getattr(self.lib, "sdpc_step_workspace_init", None)
        if init is not None:                # arm the z-buffers once: every share call re-arms what it used (no per-step memset)
            cabi.check(self.lib, init(_ptr(self.workspace), nbytes, B, H, W, self.geo.R, self._stream()),
                       "sdpc_step_workspace_init")
        self.too_high = torch.zeros(1, dtype=torch.int32, device=device)
        # debug=True: candidate-level (row/col/valid, selects the legacy full scatter kernel) + cell-level dumps;
        # debug="cells": only the per-cell dumps, so the production (compacted, fp32-guarded) scatter runs
        self.debug = None
        self.key_shift_override = 0
        self.winner_mode = 0          # sdpc_step_params.winner_mode: 0 verify where it matters, 1 verify all, 2 exact traversal
        if debug:
            R = self.geo.R
            i32 = dict(device=device, dtype=torch.int32)
            self.debug = dict(cnt=torch.zeros(B, R, W, **i32), winner=torch.zeros(B, R, W, **i32),
                              min_d=torch.zeros(B, R, W, device=device, dtype=torch.float64))
            if debug != "cells":
                self.debug.update(row=torch.zeros(B, group_size * H * W, **i32), col=torch.zeros(B, group_size * H * W, **i32),
                                  valid=torch.zeros(B, group_size * H * W, device=device, dtype=torch.uint8))

    @staticmethod
    def _expand(t, shape, name):
        try:
            return t.expand(shape)
        except RuntimeError:
            raise ValueError(f"{name} of shape {tuple(t.shape)} does not broadcast to the sample's {tuple(shape)}") from None

    def _ws_bytes(self):
        fn = getattr(self.lib, "sdpc_step_workspace_bytes", None)
        return int(fn(self.B, self.H, self.W, self.geo.R)) if fn is not None else 256

    def params(self, step_size, noise_scale, grad_ref, corr_coef, sigma_mod, share, min_depth_filter,
               allowance, sky_filter, nan_to_num=True):
        g = self.geo
        p = cabi.StepParams()
        p.n_views, p.group_size, p.height, p.width, p.big_rows = self.B, self.A, self.H, self.W, g.R
        p.variant, p.share, p.nan_to_num, p.sky_filter = self.variant, int(share), int(nan_to_num), int(sky_filter)
        p.tgt_first, p.tgt_count = self.tgt_first, self.tgt_count
        p.scalar_div_recip = int(self.scalar_div_recip)
        p.key_shift_override = int(self.key_shift_override)
        p.winner_mode = int(self.winner_mode)
        p.step_size, p.noise_scale = float(step_size), float(noise_scale)
        p.grad_ref, p.corr_coef, p.sigma_mod = float(grad_ref), float(corr_coef), float(sigma_mod)
        p.min_depth_thr = min_depth_threshold(sigma_mod) if min_depth_filter else -1.0
        p.allowance = float(allowance) if allowance is not None else -1.0
        p.h_min, p.dh, p.big_row_min, p.dv = g.h_min, g.dh, g.big_row_min, g.dv
        return p

    def buffers(self, x, grad, noise, grad_likelihood=None, new_images=None):
        b = cabi.StepBuffers()
        b.x, b.grad, b.noise = _ptr(x), _ptr(grad), _ptr(noise)
        b.refer, b.mask, b.sky, b.exist = _ptr(self.refer), _ptr(self.mask), _ptr(self.sky), _ptr(self.exist)
        b.to_world, b.from_world, b.origins = _ptr(self.to_world), _ptr(self.from_world), _ptr(self.origins)
        g = self.geo
        b.cos_az, b.sin_az, b.cos_el, b.sin_el = _ptr(g.cos_az), _ptr(g.sin_az), _ptr(g.cos_el), _ptr(g.sin_el)
        b.grad_likelihood, b.new_images, b.too_high = _ptr(grad_likelihood), _ptr(new_images), _ptr(self.too_high)
        if self.debug is not None:
            d = self.debug
            if "row" in d:
                b.dbg_row, b.dbg_col, b.dbg_valid = _ptr(d["row"]), _ptr(d["col"]), _ptr(d["valid"])
            b.dbg_cnt, b.dbg_winner, b.dbg_min_d = _ptr(d["cnt"]), _ptr(d["winner"]), _ptr(d["min_d"])
        return b

    def _stream(self):
        if torch.device(self.device).type != "cuda":       # test-only host emulation
            return None
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def step(self, p, b):
        """update (+ share when p.share) in place on b.x."""
        st = self.lib.sdpc_langevin_reproject_step(C.byref(p), C.byref(b), _ptr(self.workspace),
                                                   self.workspace.numel(), self._stream())
        cabi.check(self.lib, st, "sdpc_langevin_reproject_step")

    def step_host(self, p, b, x_host, new_images_host=None, scorenet=None, labels=None, grad_host=None, noise_host=None):
        """The whole step on HOST sample buffers (sdpc_langevin_reproject_step_host): x_host -> b.x, score forward of
        `scorenet` (an sdpc_b200 NCSN_LiDAR_small) into b.grad, update (+ share), b.x -> x_host, newImages -> host.
        Everything is enqueued on the current stream; the caller synchronises before reading the host buffers."""
        handle = ws = None
        ws_bytes = 0
        if scorenet is not None:
            handle, ws, ws_bytes = scorenet.handle_and_workspace(self.B, self.H, self.W, self.device)
        st = self.lib.sdpc_langevin_reproject_step_host(
            C.byref(p), C.byref(b), handle, _ptr(labels), ws, ws_bytes, _ptr(x_host), _ptr(grad_host), _ptr(noise_host),
            _ptr(new_images_host), _ptr(self.workspace), self.workspace.numel(), self._stream())
        cabi.check(self.lib, st, "sdpc_langevin_reproject_step_host")

    def kernel_launches(self, p, b):
        """kernels one step() call with these parameters launches (bench.py's gpu_launches)."""
        n = self.lib.sdpc_step_kernel_launches(C.byref(p), C.byref(b))
        cabi.check(self.lib, min(n, 0), "sdpc_step_kernel_launches")
        return int(n)

    def update_only(self, p, b):
        st = self.lib.sdpc_langevin_update(C.byref(p), C.byref(b), _ptr(self.workspace), self.workspace.numel(),
                                           self._stream())
        cabi.check(self.lib, st, "sdpc_langevin_update")

    def share_only(self, p, b):
        st = self.lib.sdpc_crossview_share(C.byref(p), C.byref(b), _ptr(self.workspace), self.workspace.numel(),
                                           self._stream())
        cabi.check(self.lib, st, "sdpc_crossview_share")

    def local_max(self):
        """float32 [1] device tensor: max |x0| over this rank's views after the last update."""
        out = torch.empty(1, dtype=torch.float32, device=self.device)
        cabi.check(self.lib, self.lib.sdpc_step_read_max(_ptr(self.workspace), _ptr(out), self._stream()), "read_max")
        return out

    def merge_max(self, others):
        cabi.check(self.lib, self.lib.sdpc_step_merge_max(_ptr(self.workspace), _ptr(others), others.numel(),
                                                          self._stream()), "merge_max")
