"""point_cloud_to_range_image: host mirror of LiDARGen/datasets/lidar_utils.py:54-347 on the C ABI
(row N1 of SURVEY.md 8f).  Same arguments and return tuple as the reference (numpy in, numpy out); the projection,
z-buffer, flips and the obfuscation scan run on the GPU.  CUDA only: no CPU fallback."""
import ctypes as C
import math

import numpy as np
import torch

from . import cabi


def project_device(pc, origin, return_remission, H, W):
    """pc: float64 [N, >=3(+1)] CUDA tensor.  Returns device tensors (depth f64, intensity f64 or None, obfuscation u8,
    sky u8, index f64), all [H, W] and flipped like the reference's outputs."""
    lib = cabi.load()
    dev = pc.device
    p = cabi.ProjectionParams()
    p.n_points, p.point_stride = pc.shape[0], pc.shape[1]
    p.intensity_col = 3 if return_remission else -1
    p.height, p.width = H, W
    o = np.asarray(origin, dtype=np.float64).reshape(-1)
    p.origin[0], p.origin[1], p.origin[2] = float(o[0]), float(o[1]), float(o[2])
    # lidar_utils.py:94-118: LiDARGen's KITTI specs (360 x 28 degrees, +3 degrees above the horizon)
    p.dh = math.radians(360) / W
    p.dv = math.radians(28) / H
    p.h_min = W // (-2) * p.dh + p.dh / 2
    p.v_min = math.radians(3 - 28)
    f64 = dict(dtype=torch.float64, device=dev)
    depth, inten, index = torch.empty(H, W, **f64), torch.zeros(H, W, **f64), torch.empty(H, W, **f64)
    obf = torch.empty(H, W, dtype=torch.uint8, device=dev)
    sky = torch.empty(H, W, dtype=torch.uint8, device=dev)
    ws = torch.empty(int(lib.sdpc_projection_workspace_bytes(H, W)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        st = lib.sdpc_pointcloud_to_range_image(C.byref(p), C.c_void_p(pc.data_ptr()), C.c_void_p(depth.data_ptr()),
                                                C.c_void_p(inten.data_ptr()) if return_remission else None,
                                                C.c_void_p(obf.data_ptr()), C.c_void_p(sky.data_ptr()),
                                                C.c_void_p(index.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(), stream)
        cabi.check(lib, st, "sdpc_pointcloud_to_range_image")
    return depth, (inten if return_remission else None), obf, sky, index


def point_cloud_to_range_image(point_cloud, origin, return_remission=False, return_points=False,
                               provided_origin=False, rowMax=64, colMax=1024, saveNum=0, device="cuda"):
    if not torch.cuda.is_available():
        raise cabi.SdpcError("point_cloud_to_range_image needs a CUDA device: there is no CPU fallback")
    dev = torch.device(device)
    pc = torch.as_tensor(np.ascontiguousarray(point_cloud, dtype=np.float64)).to(dev)
    depth, inten, obf, sky, index = project_device(pc, origin, return_remission, int(rowMax), int(colMax))
    d, ob, sk, ix = depth.cpu().numpy(), obf.cpu().numpy().astype(bool), sky.cpu().numpy().astype(bool), index.cpu().numpy()
    if return_remission:
        return d, inten.cpu().numpy(), ob, saveNum, sk, ix
    return d, ob, saveNum, sk, ix
