"""Multi-view dataset assembly on the C ABI (row N2 of SURVEY.md 8f): the per-view work of the reference's KITTI-360
datasets (LiDARGen/datasets/kitti360_im_8Batch.py:49-68, 94-304; kitti360_im_AllForOne.py and
kitti360_im_simultenous_densification.py pick other poses / origins but run the same steps).

* `velo_to_world_poses`  -- the calibration chain velo -> cam -> pose -> world (:49-68), float64 on the host (a handful of
  4x4 products per drive, exactly the reference's expressions).
* `assemble_view`        -- one dataset item: the raw scan is moved into the target frame's sensor coordinates, projected to
  a range image (row N1 kernels) and post-processed into the sampler's inputs; the target frame's own scan gives the ground
  truth.  The points stay on the GPU from the .bin bytes to the finished images.
* `assemble_densification_view` -- the densification dataset's item: no pose change, the scan is thinned to the points
  that own a pixel outside the blanked quarter of the columns, then rendered from the view's origin.
* `KITTI360Line` / `KITTI360AllForOne` / `KITTI360Densification` -- file-backed `Dataset`s with the reference's item layout
  for Line.yml / Inpainting.yml / Densification.yml (root path as an argument instead of the hard-coded /data/KITTI-360).
* `ItemBatches`          -- consecutive items collated into the batch tuple the runners unpack (`b200.data_root` in the yml).

CUDA only: there is no CPU fallback."""
import ctypes as C
import os

import numpy as np
import torch

from . import cabi
from .lidar_utils import project_device

MAX_RANGE = 2057.701                     # kitti360_im_8Batch.py:184


def _h(m34):
    return np.concatenate((np.reshape(m34, [3, 4]), np.array([0., 0., 0., 1.]).reshape(1, 4)))


def velo_to_world_poses(cam_to_velo, cam_to_pose, poses):
    """cam_to_velo: 12 values (calib_cam_to_velo.txt); cam_to_pose: 12 values (first row of calib_cam_to_pose.txt);
    poses: [F, 13] rows of poses.txt (frame, 3x4).  Returns (frames - 1, {frame: velo -> world 4x4})."""
    velo_to_cam = np.linalg.inv(_h(cam_to_velo))
    velo_to_pose = np.matmul(_h(cam_to_pose), velo_to_cam)
    poses = np.asarray(poses, dtype=np.float64)
    frames = poses[:, 0] - 1
    table = {}
    for frame, pose in zip(frames, np.reshape(poses[:, 1:], [-1, 3, 4])):
        table[frame] = np.matmul(_h(pose), velo_to_pose)
    return frames, table


def _postprocess(lib, dev, depth, inten, obf, sky, H, W, want_masks):
    ch = 2 if inten is not None else 1
    real = torch.empty(ch, H, W, dtype=torch.float64, device=dev)
    known = torch.empty(ch, H, W, dtype=torch.uint8, device=dev) if want_masks else None
    notsky = torch.empty(1, H, W, dtype=torch.uint8, device=dev) if want_masks else None
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    st = lib.sdpc_range_image_postprocess(ptr(depth), ptr(inten), ptr(obf) if want_masks else None, ptr(sky) if want_masks else None,
                                          H, W, MAX_RANGE, ptr(real), ptr(known), ptr(notsky), stream)
    cabi.check(lib, st, "sdpc_range_image_postprocess")
    return real, known, notsky


def assemble_view(scan, goal_scan, to_world_src, to_world_dst, origin=None, return_remission=True, rowMax=64, colMax=1024,
                  device="cuda"):
    """scan, goal_scan: float32 [N,4] raw Velodyne points of the source frame and of the target frame; to_world_src /
    to_world_dst: their velo -> world matrices.  Returns the reference's item without the frame number:
    (real [C,H,W] f64, known mask [C,H,W] bool, not-sky [1,H,W] bool, index [1,H,W] f64, toWorld [1,4,4], fromWorld [1,4,4],
    goalDepth [C,H,W] f64, toOGView [4,4])."""
    lib = cabi.load()
    if not torch.cuda.is_available():
        raise cabi.SdpcError("assemble_view needs a CUDA device: there is no CPU fallback")
    dev = torch.device(device)
    H, W = int(rowMax), int(colMax)
    to_src = np.ascontiguousarray(to_world_src, dtype=np.float64)
    to_dst = np.ascontiguousarray(to_world_dst, dtype=np.float64)
    to_og_view = np.linalg.inv(to_src)
    from_world = np.linalg.inv(to_dst)
    origin = np.zeros(3) if origin is None else np.asarray(origin, dtype=np.float64)
    dptr = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        raw = torch.as_tensor(np.ascontiguousarray(scan, dtype=np.float32)).to(dev)
        moved = torch.empty(raw.shape[0], 4, dtype=torch.float64, device=dev)
        cabi.check(lib, lib.sdpc_transform_scan(C.c_void_p(raw.data_ptr()), raw.shape[0], dptr(to_src), dptr(from_world),
                                                C.c_void_p(moved.data_ptr()), stream), "sdpc_transform_scan")
        depth, inten, obf, sky, index = project_device(moved, origin, return_remission, H, W)
        real, known, notsky = _postprocess(lib, dev, depth, inten, obf, sky, H, W, True)
        goal = torch.as_tensor(np.ascontiguousarray(goal_scan, dtype=np.float64)).to(dev)
        gdepth, ginten, _, _, _ = project_device(goal, origin, return_remission, H, W)
        goal_real, _, _ = _postprocess(lib, dev, gdepth, ginten, None, None, H, W, False)
    return (real.cpu().numpy(), known.cpu().numpy().astype(bool), notsky.cpu().numpy().astype(bool),
            index.cpu().numpy()[None], to_dst[None], from_world[None], goal_real.cpu().numpy(), to_og_view)


def assemble_densification_view(scan, to_world, modifications, view, return_remission=True, rowMax=64, colMax=1024,
                                device="cuda"):
    """One item of `KITTI360_im_simultaneous_densification` (kitti360_im_simultenous_densification.py, `__getitem__`).
    scan: float32 [N,4] raw points of the frame, to_world: its velo -> world matrix, modifications: [>=A,3] origins.
    The scan is rendered from `modifications[0]`, the first colMax // 4 columns are blanked and only the points that own
    one of the remaining pixels survive, in row-major pixel order (`scanPoints[index[index >= 0]]`); the thinned scan is
    rendered from `modifications[view]`, the full scan from the same origin is the ground truth.  View 0 replaces its
    unknown-pixel mask by the blanked quarter alone.  Returns the tuple of `assemble_view`."""
    lib = cabi.load()
    if not torch.cuda.is_available():
        raise cabi.SdpcError("assemble_densification_view needs a CUDA device: there is no CPU fallback")
    dev = torch.device(device)
    H, W = int(rowMax), int(colMax)
    mods = np.asarray(modifications, dtype=np.float64)
    to_world = np.ascontiguousarray(to_world, dtype=np.float64)
    to_og_view = np.linalg.inv(to_world)
    with torch.cuda.device(dev):
        full = torch.as_tensor(np.ascontiguousarray(scan, dtype=np.float64)).to(dev)
        _, _, _, _, owner = project_device(full, mods[0], return_remission, H, W)
        owner[:, :W // 4] = -2
        thinned = full.index_select(0, owner[owner >= 0].long()).contiguous()
        depth, inten, obf, sky, index = project_device(thinned, mods[view], return_remission, H, W)
        real, known, notsky = _postprocess(lib, dev, depth, inten, obf, sky, H, W, True)
        gdepth, ginten, _, _, _ = project_device(full, mods[view], return_remission, H, W)
        goal_real, _, _ = _postprocess(lib, dev, gdepth, ginten, None, None, H, W, False)
    known = known.cpu().numpy().astype(bool)
    if view == 0:
        known = np.ones_like(known)
        known[:, :, :W // 4] = False
    return (real.cpu().numpy(), known, notsky.cpu().numpy().astype(bool), index.cpu().numpy()[None], to_world[None],
            to_og_view[None], goal_real.cpu().numpy(), to_og_view)


class KITTI360Line(torch.utils.data.Dataset):
    """Line configuration (`KITTI360_im_8batch`, kitti360_im_8Batch.py): item idx = (frame idx // A, view idx % A); view k
    re-renders the frame's scan from the pose 5 * (k + 1) frames ahead and pairs it with that frame's own scan."""

    def __init__(self, root, config, drive="2013_05_28_drive_0000_sync", device="cuda"):
        self.root, self.drive, self.device = root, drive, device
        self.return_remission = config.data.channels == 2
        self.batchSize = config.sampling.actualBatchSize
        self.rowMax, self.colMax = config.data.image_size, config.data.image_width
        calib = os.path.join(root, "calibration")
        self.frames, self.Tr_pose_world = velo_to_world_poses(
            np.loadtxt(os.path.join(calib, "calib_cam_to_velo.txt")),
            np.loadtxt(os.path.join(calib, "calib_cam_to_pose.txt"))[0],
            np.loadtxt(os.path.join(root, "data_poses", drive, "poses.txt")))
        self.length = len(self.frames) * self.batchSize

    def __len__(self):
        return self.length

    def scan_path(self, frame):
        return os.path.join(self.root, "data_3d_raw", self.drive, "velodyne_points", "data", str(int(frame)).zfill(10) + ".bin")

    def load_scan(self, frame):
        path = self.scan_path(frame)
        if not os.path.isfile(path):
            raise RuntimeError('%s does not exist!' % path)
        return np.reshape(np.fromfile(path, dtype=np.float32), [-1, 4])

    def __getitem__(self, idx):
        view, pose_num = idx % self.batchSize, idx // self.batchSize
        frame = self.frames[pose_num]
        wanted = min(pose_num + (view + 1) * 5, len(self.frames) - 1)          # movementModifier = 5 (:146-152)
        item = assemble_view(self.load_scan(frame), self.load_scan(self.frames[wanted]), self.Tr_pose_world[frame],
                             self.Tr_pose_world[self.frames[wanted]], None, self.return_remission, self.rowMax,
                             self.colMax, self.device)
        return item + (int(frame),)


class KITTI360AllForOne(KITTI360Line):
    """Inpainting configuration (`KITTI360_im_AllForOne`, kitti360_im_AllForOne.py:94-355): every view of a group re-renders
    the frame's scan from the pose 2 * 5 frames ahead; view k looks at it from the origin `config.data.modifications[k]`."""

    def __init__(self, root, config, drive="2013_05_28_drive_0000_sync", device="cuda"):
        super().__init__(root, config, drive, device)
        self.modifications = np.array(config.data.modifications)

    def __getitem__(self, idx):
        view, pose_num = idx % self.batchSize, idx // self.batchSize
        frame = self.frames[pose_num]
        wanted = min(pose_num + 2 * 5, len(self.frames) - 1)                   # goalMovement * movementModifier (:160-171)
        item = assemble_view(self.load_scan(frame), self.load_scan(self.frames[wanted]), self.Tr_pose_world[frame],
                             self.Tr_pose_world[self.frames[wanted]], self.modifications[view], self.return_remission,
                             self.rowMax, self.colMax, self.device)
        return item + (int(frame),)


class KITTI360Densification(KITTI360AllForOne):
    """Densification configuration (`KITTI360_im_simultaneous_densification`): the frame's own scan, thinned, seen from
    `config.data.modifications[k]`; see `assemble_densification_view`."""

    def __getitem__(self, idx):
        view, pose_num = idx % self.batchSize, idx // self.batchSize
        frame = self.frames[pose_num]
        item = assemble_densification_view(self.load_scan(frame), self.Tr_pose_world[frame], self.modifications, view,
                                           self.return_remission, self.rowMax, self.colMax, self.device)
        return item + (int(frame),)


class ItemBatches:
    """`batch(i)`: items [i*B, (i+1)*B) of a multi-view dataset stacked the way the DataLoader's default collate hands them
    to the runners (ncsn_runner_kitti_simultaneous.py:543-556 unpacks exactly this tuple).  Batches are walked in order;
    the reference's `MySampler` shuffles the batch order even with random=False (SURVEY.md 8a quirk ix)."""

    def __init__(self, dataset, batch_size):
        self.dataset, self.batch_size = dataset, int(batch_size)

    def __len__(self):
        return len(self.dataset) // self.batch_size

    def batch(self, index):
        if not 0 <= index < len(self):
            raise IndexError(f"batch {index} outside [0, {len(self)})")
        items = [self.dataset[i] for i in range(index * self.batch_size, (index + 1) * self.batch_size)]
        cols = list(zip(*items))
        out = [torch.from_numpy(np.stack([np.asarray(v) for v in c])) for c in cols[:-1]]
        return tuple(out) + (torch.as_tensor(np.asarray(cols[-1])),)
