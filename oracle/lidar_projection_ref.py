"""Point cloud -> range image oracle (test infrastructure, see oracle/__init__.py).

numpy restatement of /root/reference/LiDARGen/datasets/lidar_utils.py:54-347 (point_cloud_to_range_image):
spherical projection with clamp-to-edge, nearest-depth z-buffer per pixel, 180-degree flip of every image,
then the row-sequential obfuscation scan.  Quirks kept: row 0 / column 0 never receive points (`> 0` test after
the clamp, :164-166), a nearest depth of exactly 0 reads as empty (:183), the sky mask is computed and then
cleared (:295).
"""
import math

import numpy as np

MAX_RANGE = 2057.701


def projection_constants(H, W):
    dh = math.radians(360) / W
    dv = math.radians(28) / H
    h_min = W // (-2) * dh + dh / 2
    v_min = math.radians(3 - 28)
    return dh, dv, h_min, v_min


def project_points(xyz, origin, H, W):
    """per point: depth, planar range, clamped (row, col) and the in-grid flag (lidar_utils.py:150-166)."""
    dh, dv, h_min, v_min = projection_constants(H, W)
    rel = xyz - origin
    xy2 = np.square(rel[:, 0]) + np.square(rel[:, 1])
    depth = np.sqrt(xy2 + np.square(rel[:, 2]))
    horiz = np.arctan2(rel[:, 1], rel[:, 0])
    xy = np.sqrt(xy2)
    vert = np.arctan2(rel[:, 2], xy)
    col = np.round(np.divide(horiz - h_min, dh)).astype(int)
    row = np.round(np.divide(vert - v_min, dv)).astype(int)
    col = np.clip(col, 0, W - 1).astype(np.int32)
    row = np.clip(row, 0, H - 1).astype(np.int32)
    in_grid = (col > 0) & (col < W) & (row > 0) & (row < H)
    return depth, xy, row, col, in_grid


def point_cloud_to_range_image(point_cloud, origin, return_remission=False, rowMax=64, colMax=1024):
    """returns dict(depth, intensity, obfuscation, sky, index, xy, winner) - images already flipped."""
    H, W = rowMax, colMax
    pc = np.asarray(point_cloud)
    depth, xy, row, col, ok = project_points(pc[:, :3], np.asarray(origin), H, W)
    n = len(pc)
    pix = row.astype(np.int64) * W + col
    cand = np.flatnonzero(ok)
    order = cand[np.lexsort((cand, depth[cand]))]            # nearest first, ties -> smallest point index
    first = np.unique(pix[order], return_index=True)[1]
    win = order[first]
    winner = np.full(H * W, -1, dtype=np.int64)
    winner[pix[win]] = win
    has = (winner >= 0) & (np.where(winner >= 0, depth[np.maximum(winner, 0)], 0.0) != 0)
    safe = np.maximum(winner, 0)
    img_depth = np.where(has, depth[safe], MAX_RANGE).reshape(H, W)
    img_xy = np.where(has, xy[safe], MAX_RANGE).reshape(H, W)
    img_idx = np.where(has, safe.astype(np.float64), -1.0).reshape(H, W)
    img_int = np.zeros((H, W))
    if return_remission:
        img_int = np.where(has, pc[safe, 3].astype(np.float64), 0.0).reshape(H, W)
    img_depth, img_xy, img_idx, img_int = (np.flip(a).copy() for a in (img_depth, img_xy, img_idx, img_int))
    win_img = np.flip(np.where(has, winner, -1).reshape(H, W)).copy()
    # obfuscation scan (lidar_utils.py:262-305)
    obf = np.zeros((H, W), dtype=bool)
    sky_prev = np.ones(W, dtype=bool)                          # rows 0 and 1 are "sky"
    min_depth = np.full(W, MAX_RANGE)
    for r in range(2, H - 1):
        obf[r] = img_xy[r] > min_depth + 5
        e = ((img_xy[r] != min_depth).astype(int) + (img_xy[r - 1] != min_depth).astype(int)
             + (img_xy[r + 1] != min_depth).astype(int))
        ep = np.concatenate(([0], e, [0]))
        eq = (ep[1:-1] + ep[:-2] + ep[2:]) <= 1
        cur_sky = eq & sky_prev
        sky_prev = cur_sky
        upd = ~cur_sky
        min_depth[upd] = np.minimum(img_xy[r], min_depth)[upd]
    obf[-1] = img_xy[-1] > min_depth + 5
    sky = np.zeros((H, W), dtype=bool)                          # the reference clears it before returning
    return dict(depth=img_depth, intensity=img_int, obfuscation=obf, sky=sky, index=img_idx, xy=img_xy, winner=win_img)
