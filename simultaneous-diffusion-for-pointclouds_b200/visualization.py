"""Output stage on the C ABI (row N3 of SURVEY.md 8f): what the reference does with a sampled range image.

* `range_image_to_pointcloud(image)` -- the numeric part of `visualize_tensor` (LiDARGen/visualization.py:12-43):
  depth = 2^(6 r) - 1, the sensor's yaw / pitch grid, xyz of the pixels with 0.5 m < depth < 63 m in row-major order.
  The open3d rendering that follows in the reference (:44-62) is presentation and out of scope.
* `depth_intensity_errors(pred, gt, inp)` -- the per-view L1 error sums of the evaluation notebook
  (MeasureResults/QuantifyingNotebookSynthesis_Line.ipynb, cell 1).

CUDA only: there is no CPU fallback."""
import ctypes as C

import numpy as np
import torch

from . import cabi


def yaw_pitch_grid(H=64, W=1024):
    """float64 yaw [W] and pitch [H] of visualization.py:20-33 (fov_up 3 deg, fov_down -25 deg)."""
    fov_up = 3.0 / 180.0 * np.pi
    fov_down = -25.0 / 180.0 * np.pi
    fov = abs(fov_down) + abs(fov_up)
    x = np.arange(0, float(W)) * (1 / float(W))
    y = np.arange(0, float(H)) * (1 / float(H))
    return np.pi * (x * 2 - 1), (1.0 - y) * fov - abs(fov_down)


def _need_cuda(what):
    if not torch.cuda.is_available():
        raise cabi.SdpcError(f"{what} needs a CUDA device: there is no CPU fallback")


def range_images_to_pointclouds(images, device="cuda", with_pixels=False):
    """images: [V,2,H,W] float32 (tensor or array).  Returns a list of V tuples (xyz float64 [M,3], intensity float32 [M])
    (+ pixel index int32 [M] when with_pixels), device tensors."""
    lib = cabi.load()
    _need_cuda("range_images_to_pointclouds")
    dev = torch.device(device)
    img = torch.as_tensor(images, dtype=torch.float32).to(dev).contiguous()
    V, ch, H, W = img.shape
    if ch != 2:
        raise ValueError("expected [V,2,H,W] range / intensity images")
    yaw, pitch = yaw_pitch_grid(H, W)
    tabs = [torch.from_numpy(np.ascontiguousarray(t)).to(dev) for t in (np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch))]
    xyz = torch.empty(V, H * W, 3, dtype=torch.float64, device=dev)
    inten = torch.empty(V, H * W, dtype=torch.float32, device=dev)
    pix = torch.empty(V, H * W, dtype=torch.int32, device=dev)
    cnt = torch.empty(V, dtype=torch.int32, device=dev)
    ws = torch.empty(int(lib.sdpc_points_workspace_bytes(V, H, W)), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        st = lib.sdpc_range_image_to_points(C.c_void_p(img.data_ptr()), V, H, W, *[C.c_void_p(t.data_ptr()) for t in tabs],
                                            C.c_void_p(xyz.data_ptr()), C.c_void_p(inten.data_ptr()),
                                            C.c_void_p(pix.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                            C.c_void_p(ws.data_ptr()), ws.numel(), stream)
        cabi.check(lib, st, "sdpc_range_image_to_points")
    n = cnt.cpu().tolist()
    out = []
    for v in range(V):
        item = (xyz[v, :n[v]], inten[v, :n[v]])
        out.append(item + (pix[v, :n[v]],) if with_pixels else item)
    return out


def range_image_to_pointcloud(image, device="cuda"):
    """image: [2,H,W] (the reference's `visualize_tensor(image)` argument).  Returns (xyz [M,3] float64, intensity [M])
    as numpy arrays: `xyz` is the reference's `pts[mask, :]`, `intensity` its `lidar_intensity.flatten()[mask]`."""
    xyz, inten = range_images_to_pointclouds(torch.as_tensor(image)[None], device)[0]
    return xyz.cpu().numpy(), inten.cpu().numpy()


ERROR_KEYS = ("depth_l1", "intensity_l1", "depth_l1_input", "intensity_l1_input", "depth_sum_input", "pixels", "input_pixels")


def depth_intensity_errors(pred, gt, inp, device="cuda"):
    """pred, gt, inp: [V,2,H,W] float32 (sample, ground truth, masked input).  Returns a dict of float64 [V] arrays, the
    per-view sums the notebook accumulates into totalDistanceError / totalIntensityError / ...Input / totalPixels."""
    lib = cabi.load()
    _need_cuda("depth_intensity_errors")
    dev = torch.device(device)
    t = [torch.as_tensor(a, dtype=torch.float32).to(dev).contiguous() for a in (pred, gt, inp)]
    V, ch, H, W = t[0].shape
    if ch != 2 or any(a.shape != t[0].shape for a in t):
        raise ValueError("expected three [V,2,H,W] tensors of the same shape")
    out = torch.empty(V, 8, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        st = lib.sdpc_depth_intensity_errors(*[C.c_void_p(a.data_ptr()) for a in t], V, H, W, C.c_void_p(out.data_ptr()), stream)
        cabi.check(lib, st, "sdpc_depth_intensity_errors")
    o = out.cpu().numpy()
    return {k: o[:, i].copy() for i, k in enumerate(ERROR_KEYS)}
