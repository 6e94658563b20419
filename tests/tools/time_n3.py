"""Row N3 timing: CUDA range image -> point cloud and error sums (device-resident, CUDA events) next to the numpy oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import sdpc_b200  # noqa
from sdpc_b200.visualization import range_images_to_pointclouds, depth_intensity_errors
from oracle import output_stage_ref as osr

V = 8
rng = np.random.default_rng(1)
img = rng.uniform(0, 1, size=(V, 2, 64, 1024)).astype(np.float32)
inp = (img * (rng.uniform(size=img.shape) < 0.6)).astype(np.float32)
pred = np.clip(img + 0.01, 0, 1).astype(np.float32)
t0 = time.perf_counter()
for v in range(V):
    osr.points_ref(img[v])
cpu_pts = (time.perf_counter() - t0) * 1e3
t0 = time.perf_counter()
osr.error_sums_ref(pred, img, inp)
cpu_err = (time.perf_counter() - t0) * 1e3
dev = torch.device("cuda:0")
d_img, d_inp, d_pred = (torch.as_tensor(a).to(dev) for a in (img, inp, pred))


def timeit(fn, it=30):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it * 1e3


pts_us = timeit(lambda: range_images_to_pointclouds(d_img))
err_us = timeit(lambda: depth_intensity_errors(d_pred, d_img, d_inp))
nbytes = V * 65536 * (8 + 28)            # read r + intensity, write xyz (24 B) + intensity (4 B) per kept pixel at most
print(f"N3 {V} views: points numpy {cpu_pts:.1f} ms | CUDA {pts_us:.0f} us per call incl. host mirror (<= {nbytes / 1e6:.1f} MB moved); "
      f"error sums numpy {cpu_err:.1f} ms | CUDA {err_us:.0f} us")
