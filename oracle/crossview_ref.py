"""Cross-view consistency oracle: un-project -> pose -> re-project -> z-buffer -> fusion.

TEST INFRASTRUCTURE (see oracle/__init__.py).  A semantic restatement (flat
scatter/segment reductions instead of the reference's sort / unique_consecutive /
sparse_coo pipeline) of

  * pose-matrix variant      /root/reference/LiDARGen/models/KITTISampling.py:29-102,160-430
  * translation-only variant /root/reference/LiDARGen/models/__init__.py:134-175,224-231,255-520

Written with device-agnostic torch ops in the reference's dtypes (fp32 range
decode, float64 geometry, int32 indices) so that on the CPU it reproduces the
reference's indices bit-for-bit, and on a CUDA device it evaluates the same
libm (powf / atan2 / log2) the CUDA kernels use.
"""
import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class Geometry:
    """Sensor-model constants (KITTISampling.py:29-78,101-102)."""
    H: int
    W: int
    R: int                 # bigRowCount
    dh: float              # horizontalAngles
    dv: float              # verticalAngles
    h_min: float           # horizontalMin
    v_min: float           # verticalMin
    big_row_min: float     # bigRowMin
    azimuth: torch.Tensor   # [W] float64
    elevation: torch.Tensor  # [H] float64


def make_geometry(H, W, device="cpu"):
    h_scope_min, h_scope_max = -180, 180
    v_scope_max, v_scope_min = 3, -25
    h_scope = h_scope_max - h_scope_min
    v_scope = v_scope_max - v_scope_min
    dh = math.radians(h_scope) / W
    dv = math.radians(v_scope) / H
    h_min = ((W * h_scope_min) // h_scope) * dh + dh / 2
    R = int((max(abs(v_scope_min), abs(v_scope_max)) * 2) * H // v_scope)
    big_row_min = (R // -2) * dv + dv / 2
    v_min = ((H * v_scope_min) // v_scope) * dv + dv / 2
    az = torch.from_numpy(np.arange(W - 1, -1, -1) * dh + h_min).to(device)
    el = torch.from_numpy(np.arange(H - 1, -1, -1) * dv + v_min).to(device)
    return Geometry(H, W, R, dh, dv, h_min, v_min, big_row_min, az, el)


def translation_origins(modification_list):
    """a-5's originList (models/__init__.py:224-231): a fp32 log2/pow round trip
    that collapses every configured offset to sign(m)*10 (10.00000095 for |m|=5)."""
    og = modification_list.unsqueeze(-1).unsqueeze(-1)
    o = torch.log2(torch.abs(og) + 1) / 6
    o = torch.pow(2, o * 6) - 1
    return o / (og + 0.00000001) * 10      # [M,3,1,1] float32


def decode_range(x0, sigma_mod):
    """fp32 log-range -> signed metres (KITTISampling.py:161-166)."""
    is_neg = x0 < 0
    sign = torch.ones_like(x0) - is_neg.int() * 2
    dist = (torch.pow(2, torch.abs(x0) * 6 / sigma_mod) - 1) * sign
    return dist, is_neg


def min_depth_threshold(sigma_mod):
    """fp32 threshold of the 'setting 5' / a-5 min-depth filter (KITTISampling.py:273-274)."""
    return torch.log2(torch.tensor(0.2) + 1) / 6 * sigma_mod


@torch.no_grad()
def project_candidates(x, geo, sigma_mod, A, to_world=None, from_world=None, origins=None, targets=None):
    """Steps 3-4 of SURVEY 8(a).  Returns per (target view t, source point j of t's group):
       nd [B, A*HW] float64, row/col [B, A*HW] int32 (row in the R-row grid),
       plus is_neg [B,H,W] bool and too_high (0-dim bool).
       `targets` = (first, count) restricts the TARGET views to a range inside one group (the per-rank share of a
       view-sharded step): the leading dimension of nd/row/col/is_neg is then `count`."""
    B, _, H, W = x.shape
    HW = H * W
    G = B // A
    x0 = x[:, 0]
    dist, is_neg = decode_range(x0, sigma_mod)
    too_high = torch.max(torch.abs(x0)) * 6 / sigma_mod > 50
    ca, sa = torch.cos(geo.azimuth).view(1, 1, W), torch.sin(geo.azimuth).view(1, 1, W)
    ce, se = torch.cos(geo.elevation).view(1, H, 1), torch.sin(geo.elevation).view(1, H, 1)
    px = dist * ca * ce
    py = dist * sa * ce
    pz = dist * se
    if origins is None:
        # pose-matrix variant: world = toWorld[v] . p ; rel = fromWorld[t] . world
        P = torch.stack((px.reshape(B, HW), py.reshape(B, HW), pz.reshape(B, HW),
                         torch.ones(B, HW, dtype=px.dtype, device=x.device)), 1)
        Pw = torch.bmm(to_world, P)                                  # [B,4,HW]
        cloud = Pw.view(G, A, 4, HW).permute(0, 2, 1, 3).reshape(G, 4, A * HW)
        if targets is not None:
            t0, tn = targets
            assert t0 // A == (t0 + tn - 1) // A, "a target range must stay inside one group"
            rel = torch.bmm(from_world[t0:t0 + tn], cloud[t0 // A].unsqueeze(0).expand(tn, 4, A * HW))[:, :3]
        else:
            rel = torch.cat([torch.bmm(from_world[g * A:(g + 1) * A],
                                       cloud[g].unsqueeze(0).expand(A, 4, A * HW))[:, :3]
                             for g in range(G)], 0)                  # [B,3,A*HW]
    else:
        o = origins[:A]                                              # [A,3,1,1] fp32
        rel_groups = []
        for g in range(G):
            sl = slice(g * A, (g + 1) * A)
            wx = (px[sl] + o[:, 0]).reshape(1, A * HW)
            wy = (py[sl] + o[:, 1]).reshape(1, A * HW)
            wz = (pz[sl] + o[:, 2]).reshape(1, A * HW)
            cloud = torch.stack((wx, wy, wz), 1).expand(A, 3, A * HW)
            rel_groups.append(cloud - o[:, :, 0])
        rel = torch.cat(rel_groups, 0)
        if targets is not None:
            rel = rel[targets[0]:targets[0] + targets[1]]
    xy = torch.square(rel[:, 0]) + torch.square(rel[:, 1])
    nd = torch.log2(torch.sqrt(xy + torch.square(rel[:, 2])) + 1) / 6 * sigma_mod
    horiz = torch.atan2(rel[:, 1], rel[:, 0])
    vert = torch.atan2(rel[:, 2], torch.sqrt(xy))
    colr = torch.round((horiz - geo.h_min) / geo.dh).int()
    rowr = torch.round((vert - geo.big_row_min) / geo.dv).int()
    col = colr * -1 + W - 1
    row = rowr * -1 + geo.R - 1
    if targets is not None:
        is_neg = is_neg[targets[0]:targets[0] + targets[1]]
    return nd, row, col, is_neg, too_high


@torch.no_grad()
def shared_images(x, geo, sigma_mod, A, exist_mask, sky=None, to_world=None, from_world=None,
                  origins=None, min_depth_filter=True, controlled_average=True, allowance=10.0,
                  sky_filter=False, return_debug=False, targets=None):
    """Steps 3-7 of SURVEY 8(a): the shared re-projection of every view.

    x [B,2,H,W] fp32 (post-Langevin sample); exist_mask bool [>=A,H,W]; sky bool [B,1,H,W].
    Returns new_images [B,2,H,W] fp32, image_mask [B,H,W] bool (already AND existMask[0]),
    too_high; with return_debug also a dict of per-candidate and per-pixel internals.
    `targets` = (first, count): only these target views (inside one group) are built - every output then has `count`
    leading entries; the sources are still all A views of that group.
    """
    Bx, _, H, W = x.shape
    HW, R = H * W, geo.R
    dev = x.device
    nd, row, col, is_neg, too_high = project_candidates(x, geo, sigma_mod, A, to_world, from_world, origins, targets)
    t0, B = (0, Bx) if targets is None else targets                   # B: number of target views from here on
    valid = (col > -1) & (col < W) & (row > -1) & (row < R)
    if sky_filter:            # a-5 only: source pixel's sky flag (models/__init__.py:352-355)
        G = Bx // A
        valid &= sky.reshape(G, 1, A * HW).expand(G, A, A * HW).reshape(Bx, A * HW)[t0:t0 + B]
    valid &= exist_mask[:A].reshape(1, A * HW)
    if min_depth_filter:
        valid &= nd > min_depth_threshold(sigma_mod).to(dev)

    t_idx = torch.arange(B, device=dev).view(B, 1).expand(B, A * HW)
    key = ((t_idx * R + row.long()) * W + col.long())[valid]           # flat pixel of target grid
    src = torch.arange(A * HW, device=dev).view(1, A * HW).expand(B, A * HW)[valid]
    tgt = t_idx[valid]
    inten_all = x[:, 1].reshape(Bx // A, A * HW)                     # group-major source intensities
    nd_v = nd[valid]
    in_v = inten_all[(tgt + t0) // A, src]

    n_pix = B * R * W
    cnt = torch.zeros(n_pix, dtype=torch.int64, device=dev).index_add_(0, key, torch.ones_like(key))
    sum_d = torch.zeros(n_pix, dtype=torch.float64, device=dev).index_add_(0, key, nd_v)
    sum_i = torch.zeros(n_pix, dtype=torch.float32, device=dev).index_add_(0, key, in_v)
    min_d = torch.full((n_pix,), float("inf"), dtype=torch.float64, device=dev)
    min_d.scatter_reduce_(0, key, nd_v, reduce="amin")
    is_win = nd_v == min_d[key]
    big = A * HW
    winner = torch.full((n_pix,), big, dtype=torch.int64, device=dev)
    winner.scatter_reduce_(0, key[is_win], src[is_win], reduce="amin")   # ties -> smallest source id
    n_tied = torch.zeros(n_pix, dtype=torch.int64, device=dev).index_add_(0, key[is_win], torch.ones_like(key[is_win]))
    has = cnt > 0
    win_safe = torch.where(has, winner, torch.zeros_like(winner))
    grp = (torch.arange(n_pix, device=dev) // (R * W) + t0) // A
    min_i = torch.where(has, inten_all[grp, win_safe], torch.zeros((), dtype=torch.float32, device=dev))
    min_d = torch.where(has, min_d, torch.zeros((), dtype=torch.float64, device=dev))

    scaling = cnt.to(torch.float32) + 0.000000001                       # fp32, as in the reference
    avg_d = sum_d / scaling                                             # float64
    avg_i = sum_i / scaling                                             # float32
    if controlled_average:
        m_avg = torch.pow(2, torch.abs(avg_d) * 6 / sigma_mod) - 1
        m_min = torch.pow(2, torch.abs(min_d) * 6 / sigma_mod) - 1
        far = m_avg > m_min + allowance
        avg_i = torch.where(far, min_i, avg_i)
        m_avg = torch.where(far, m_min + allowance / 5, m_avg)
        avg_d = torch.log2(m_avg + 1) / 6 * sigma_mod

    # crop rows [R-H, R) and mirror for negative-range pixels (KITTISampling.py:401-403)
    gd = avg_d.view(B, R, W)
    gi = avg_i.view(B, R, W)
    gm = has.view(B, R, W)
    top = R - H

    def crop_mirror(g, negate):
        direct = g[:, top:]
        mirrored = torch.flip(torch.roll(g, W // 2, dims=2), dims=(1,))[:, top:]
        if negate:
            mirrored = mirrored * -1
        return torch.where(is_neg, mirrored, direct)

    depth = crop_mirror(gd, True)
    inten = crop_mirror(gi, False)
    image_mask = crop_mirror(gm, False) & exist_mask[0].view(1, H, W)
    new_images = torch.stack((depth.float(), inten.float()), 1)
    if not return_debug:
        return new_images, image_mask, too_high
    dbg = dict(nd=nd, row=row, col=col, valid=valid, cnt=cnt.view(B, R, W), sum_d=sum_d.view(B, R, W),
               sum_i=sum_i.view(B, R, W), min_d=min_d.view(B, R, W), min_i=min_i.view(B, R, W),
               winner=torch.where(has, winner, torch.full_like(winner, -1)).view(B, R, W),
               n_tied=n_tied.view(B, R, W), is_neg=is_neg)
    return new_images, image_mask, too_high, dbg


@torch.no_grad()
def apply_correction(x, new_images, image_mask, sky, mask, too_high, coef):
    """Step 8 (KITTISampling.py:427-430,490)."""
    m = (image_mask.unsqueeze(1) & sky).int()
    corr = -m * torch.logical_not(mask).int() * (x - new_images)
    corr = torch.where(too_high, torch.tensor(0, device=x.device).float(), corr)
    return x + coef * corr
