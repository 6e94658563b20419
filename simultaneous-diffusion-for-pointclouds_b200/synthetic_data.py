"""Synthetic stand-in for the KITTI-360 multi-view datasets.

Yields the tuple the reference datasets return (datasets/kitti360_im_8Batch.py:304,
kitti360_im_AllForOne.py, kitti360_im_simultenous_densification.py):
  (real [B,2,H,W] f64, mask bool [B,2,H,W], sky bool [B,1,H,W], indices [B,1,H,W], toWorld [B,1,4,4] f64,
   fromWorld [B,1,4,4] f64, goalImages [B,2,H,W], toOGView [B,4,4], saveNum [B])
for batches of `batch_size` views in groups of `group` poses.  No KITTI-360 data exists offline; geometry is a
ground plane plus a few vertical walls ray-cast per view, so that neighbouring views really overlap."""
import math

import numpy as np
import torch


def _ray_dirs(H, W):
    dh = math.radians(360) / W
    dv = math.radians(28) / H
    h_min = ((W * -180) // 360) * dh + dh / 2
    v_min = ((H * -25) // 28) * dv + dv / 2
    az = np.arange(W - 1, -1, -1) * dh + h_min
    el = np.arange(H - 1, -1, -1) * dv + v_min
    ca, sa, ce, se = np.cos(az)[None, :], np.sin(az)[None, :], np.cos(el)[:, None], np.sin(el)[:, None]
    return np.stack([ca * ce, sa * ce, np.broadcast_to(se, (H, W))], -1)          # [H,W,3] sensor frame


def render_view(H, W, pose, rng, walls):
    """range image of a simple scene seen from `pose` (4x4 sensor->world)."""
    d = _ray_dirs(H, W) @ pose[:3, :3].T
    o = pose[:3, 3]
    dist = np.full((H, W), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        t = (-1.73 - o[2]) / d[..., 2]                                             # ground plane z = -1.73 m
        dist = np.where((t > 0.5), np.minimum(dist, t), dist)
        for (nx, ny, c) in walls:                                                   # vertical planes nx*x + ny*y = c
            t = (c - nx * o[0] - ny * o[1]) / (nx * d[..., 0] + ny * d[..., 1])
            dist = np.where((t > 0.5), np.minimum(dist, t), dist)
    known = np.isfinite(dist) & (dist < 63.0)
    depth = np.where(known, np.log2(np.where(known, dist, 0.0) + 1) / 6, 0.0)
    inten = np.where(known, 0.25 + 0.2 * np.sin(dist) ** 2, 0.0)
    return np.stack([depth, inten]), known


class SyntheticMultiView:
    def __init__(self, H, W, batch_size, group, mode="line", seed=1234, densify_rows=4):
        self.H, self.W, self.B, self.A, self.mode, self.seed, self.densify_rows = H, W, batch_size, group, mode, seed, densify_rows

    def poses(self, g):
        out = []
        for i in range(self.A):
            a = 0.01 * i + 0.05 * g
            T = np.eye(4)
            T[:3, :3] = [[math.cos(a), -math.sin(a), 0], [math.sin(a), math.cos(a), 0], [0, 0, 1]]
            T[:3, 3] = [5.0 * (i + 1), 0.3 * i, 0.0]
            out.append(T)
        return out

    def batch(self, index):
        rng = np.random.Generator(np.random.PCG64([self.seed, index]))
        G = self.B // self.A
        real, mask, to_w = [], [], []
        for g in range(G):
            walls = [(1.0, 0.0, 40.0 + 5 * g), (0.0, 1.0, 12.0), (0.0, 1.0, -9.0), (1.0, 0.0, -25.0)]
            for T in self.poses(g):
                img, known = render_view(self.H, self.W, T, rng, walls)
                real.append(img)
                m = known & (rng.uniform(size=known.shape) < 0.9)
                if self.mode == "densification" and len(real) % self.A == 1:       # target view keeps every 4th beam
                    keep = np.zeros_like(m)
                    keep[::self.densify_rows] = True
                    m &= keep
                mask.append(np.stack([m, m]))
                to_w.append(T)
        real = torch.from_numpy(np.stack(real))
        mask = torch.from_numpy(np.stack(mask))
        to_world = torch.from_numpy(np.stack(to_w)).unsqueeze(1)
        from_world = torch.linalg.inv(to_world)
        sky = torch.ones(self.B, 1, self.H, self.W, dtype=torch.bool)               # always True
        indices = torch.arange(self.H * self.W).view(1, 1, self.H, self.W).repeat(self.B, 1, 1, 1)
        save_num = torch.arange(index * self.B, (index + 1) * self.B)
        # toOGView is a bare 4x4 per item in the reference (no expand_dims, kitti360_im_8Batch.py:299-304): [B,4,4] after collate
        return real, mask, sky, indices, to_world, from_world, real.clone(), from_world.clone().squeeze(1), save_num
