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


# ---- benchmark workloads (bench.py; BASELINE.json configs 2-5) -------------------------------------------------------
# numpy PCG64 streams only, so every rank and every box regenerates the same bytes.
INPAINTING_MODIFICATIONS = [[0, 0, 0], [5, -5, 0], [-5, -5, 0], [0, 5, 0], [-10, 10, 0], [10, 10, 0], [-10, 0, 0],
                            [10, 0, 0]]        # Inpainting.yml's seven offsets + one more


def line_poses(B, A, step=5.0, yaw=0.01, lateral=0.3):
    """B sensor poses in groups of A, each group a line along +x with a small yaw:
    (toWorld, fromWorld) float64 [B,1,4,4] like the dataset tuple (kitti360_im_8Batch.py:304)."""
    to_world = np.zeros((B, 1, 4, 4), dtype=np.float64)
    for b in range(B):
        i, g = b % A, b // A
        a = yaw * i + 0.05 * g
        T = np.eye(4)
        T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        T[:3, 3] = [step * (i + 1), lateral * i - 0.2 * g, 0.05 * i]
        to_world[b, 0] = T
    return torch.from_numpy(to_world), torch.from_numpy(np.linalg.inv(to_world))


def smooth_range_image(B, H, W, seed):
    """[B,2,H,W] float32: log-range of d ~ U(2, 60) m smoothed along rows, intensity U(0, 0.5)."""
    r = np.random.Generator(np.random.PCG64([seed, 1]))
    d = r.uniform(2.0, 60.0, size=(B, H, W))
    k = np.ones(5) / 5
    d = np.apply_along_axis(lambda v: np.convolve(np.concatenate([v[-2:], v, v[:2]]), k, mode="valid"), 2, d)
    depth = np.clip(np.log2(d + 1) / 6, 0, 1)
    inten = r.uniform(0, 0.5, size=(B, H, W))
    return torch.from_numpy(np.stack([depth, inten], 1).astype(np.float32))


def lidargen_exist_mask(H=64, W=1024):
    """The processed beam-existence mask [H,W] bool of the reference's own data file (MeasureResults/
    existTotalLiDARGenSettings.npy after the runner's threshold + erosion, ncsn_runner_kitti_simultaneous.py:527-533),
    kept bit-packed in the package (data/exist_mask_lidargen.npz, written by tests/golden/make_golden_exist.py);
    None for another image size."""
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "exist_mask_lidargen.npz")
    if (H, W) != (64, 1024) or not os.path.exists(path):
        return None
    z = np.load(path)
    return np.unpackbits(z["packed"])[: H * W].reshape(H, W).astype(bool)


def bench_group(B, A, H, W, seed, variant="line"):
    """One rank's benchmark input: B views in groups of A.

    variant "line" (config 2): pose-matrix sampler a-4, known-pixel mask Bernoulli(0.6);
    "inpainting" (config 3): translation sampler a-5, modificationList of SURVEY 8d, view 0 of a group = target with a
    Bernoulli(0.6) mask, the supplementary views know 90 % of their pixels;
    "densification" (config 4): as inpainting with view 0's known pixels = rows 0::4 (16 of 64 beams).
    existMask: the reference's own processed fixture at 64x1024, a Bernoulli(0.68) stand-in otherwise."""
    r = np.random.Generator(np.random.PCG64([seed, 7]))
    refer = smooth_range_image(B, H, W, seed)
    known = r.uniform(size=(B, 1, H, W)) < 0.6
    ex = lidargen_exist_mask(H, W)
    if ex is None:
        ex = r.uniform(size=(H, W)) < 0.68
    x0 = torch.from_numpy(r.uniform(size=(B, 2, H, W)).astype(np.float32))
    out = dict(variant=variant, sky=torch.ones(B, 1, H, W, dtype=torch.bool),
               exist=torch.from_numpy(np.ascontiguousarray(np.broadcast_to(ex, (A, H, W)))), x=x0, refer=refer)
    if variant == "line":
        out["toWorld"], out["fromWorld"] = line_poses(B, A, step=5.0, yaw=0.01)
    else:
        if A > len(INPAINTING_MODIFICATIONS):
            raise ValueError(f"group size {A} > {len(INPAINTING_MODIFICATIONS)} configured view offsets")
        supp = r.uniform(size=(B, 1, H, W)) < 0.9
        first = (np.arange(B) % A == 0).reshape(B, 1, 1, 1)
        if variant == "densification":
            rows = np.zeros((1, 1, H, 1), dtype=bool)
            rows[:, :, 0::4] = True
            known = np.broadcast_to(rows, known.shape)
        known = np.where(first, known, supp)
        out["mods"] = torch.tensor(INPAINTING_MODIFICATIONS[:A], dtype=torch.int64)
    out["mask"] = torch.from_numpy(np.ascontiguousarray(known).astype(np.int32)).repeat(1, 2, 1, 1).contiguous()
    return out
