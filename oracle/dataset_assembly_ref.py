"""ORACLE (test infrastructure, never imported by the product): CPU restatement of the reference's per-view dataset
assembly, row N2 -- LiDARGen/datasets/kitti360_im_8Batch.py:49-68 (calibration chain) and :94-304 (`__getitem__`).

Pinned: tests/golden/make_golden_n2.py executes the reference's OWN source lines of both blocks (read from
/root/reference at generation time; the class cannot be imported here because the datasets package needs h5py and
/data/KITTI-360) against synthetic calibration files and scans; tests/test_n2_dataset_assembly.py compares this
restatement with the committed fixture."""
import numpy as np

from . import lidar_projection_ref as lp

MAX_RANGE = 2057.701


def _h(m34):
    return np.concatenate((np.reshape(m34, [3, 4]), np.array([0., 0., 0., 1.]).reshape(1, 4)))


def pose_chain(cam_to_velo, cam_to_pose, poses):
    """:49-68 -> (frames - 1, {frame: velo -> world})."""
    veloToCam = np.linalg.inv(_h(cam_to_velo))
    veloToPose = np.matmul(_h(cam_to_pose), veloToCam)
    frames = poses[:, 0] - 1
    table = {}
    for frame, pose in zip(frames, np.reshape(poses[:, 1:], [-1, 3, 4])):
        table[frame] = np.matmul(_h(pose), veloToPose)
    return frames, table


def move_scan(scan, to_world, from_world):
    """:133-141, :178-181: homogeneous points through toWorld then fromWorld, remission re-attached."""
    intensity = scan[:, -1]
    pointVals = np.concatenate((np.transpose(scan[:, :-1]), np.expand_dims(np.ones_like(intensity), 0)), 0)
    pointVals = np.matmul(to_world, pointVals)
    pointVals = np.matmul(from_world, pointVals)
    return np.transpose(np.concatenate((pointVals[:-1], np.expand_dims(intensity, 0)), 0))


def postprocess(depth, intensity, mask, sky):
    """:203-287 for one projected image (mask / sky may be None for the ground-truth image)."""
    if mask is not None:
        mask = np.where(depth >= MAX_RANGE, 1, mask)
    real = np.where(depth >= MAX_RANGE, 0, depth) + 0.0001
    real = np.clip(np.log2(real + 1) / 6, 0, 1)
    if mask is not None:
        mask = np.where(intensity >= 1, 1, mask)
    inten = np.clip(np.where(intensity >= 1, 0, intensity) + 0.0001, 0, 1.0)
    out = np.concatenate((real[None], inten[None]), axis=0)
    if mask is None:
        return out, None, None
    sky = sky.copy()
    for _ in range(3):
        sky[1:] = sky[:-1].copy()
    known = np.logical_not(np.concatenate((mask[None], mask[None]), axis=0))
    return out, known, np.logical_not(sky[None])


def assemble_view(scan, goal_scan, to_world_src, to_world_dst, H, W, origin=None):
    origin = np.zeros(3) if origin is None else origin
    to_og_view = np.linalg.inv(to_world_src)
    from_world = np.linalg.inv(to_world_dst)
    moved = move_scan(scan, to_world_src, from_world)
    r = lp.point_cloud_to_range_image(moved, origin, True, H, W)
    real, known, notsky = postprocess(r["depth"], r["intensity"], r["obfuscation"], r["sky"])
    g = lp.point_cloud_to_range_image(goal_scan.astype(np.float64), origin, True, H, W)
    goal, _, _ = postprocess(g["depth"], g["intensity"], None, None)
    return dict(real=real, known=known, notsky=notsky, index=r["index"][None], toWorld=to_world_dst[None],
                fromWorld=from_world[None], goalDepth=goal, toOGView=to_og_view, moved=moved)


def allforone_selection(pose_num, n_frames):
    """kitti360_im_AllForOne.py:160-171: every view of a group re-renders its frame's scan from the pose 2 * 5 frames
    ahead (the last pose when the drive ends earlier); the view only chooses the origin, `config.data.modifications[view]`."""
    return min(pose_num + 2 * 5, n_frames - 1)


def assemble_view_densification(scan, to_world, modifications, view, H, W):
    """kitti360_im_simultenous_densification.py `__getitem__`: no pose change.  The scan is first rendered from
    `modifications[0]`, the first quarter of the columns is blanked, and only the points that own a remaining pixel are
    kept (in row-major pixel order); that thinned scan is rendered from `modifications[view]`, the full scan from the same
    origin is the ground truth.  View 0 replaces its unknown-pixel mask by the blanked quarter alone.
    Pinned by tests/golden/make_golden_n2_variants.py (the reference's own source lines)."""
    modifications = np.asarray(modifications)
    first = lp.point_cloud_to_range_image(scan, modifications[0], True, H, W)
    index = first["index"].copy()
    index[:, :(W // 4)] = -2
    thinned = scan[index[index >= 0].astype(int)]
    origin = modifications[view]
    r = lp.point_cloud_to_range_image(thinned, origin, True, H, W)
    real, known, notsky = postprocess(r["depth"], r["intensity"], r["obfuscation"], r["sky"])
    if view == 0:
        known = np.ones_like(known)
        known[:, :, :(W // 4)] = False
    g = lp.point_cloud_to_range_image(scan, origin, True, H, W)
    goal, _, _ = postprocess(g["depth"], g["intensity"], None, None)
    to_og_view = np.linalg.inv(to_world)
    return dict(real=real, known=known, notsky=notsky, index=r["index"][None], toWorld=to_world[None],
                fromWorld=to_og_view[None], goalDepth=goal, toOGView=to_og_view, thinned=thinned)
