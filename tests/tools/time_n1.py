"""Row N1 timing: CUDA point_cloud_to_range_image (device-resident points, CUDA events) next to the numpy oracle port."""
import ctypes as C, math, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import sdpc_b200  # noqa
from sdpc_b200 import cabi
from sdpc_b200.lidar_utils import point_cloud_to_range_image
from oracle import lidar_projection_ref as lp
from tests.golden import cases

n, H, W = 120000, 64, 1024
pc, origin = cases.synthetic_scan(n, 5)
t0 = time.perf_counter()
for _ in range(3):
    lp.point_cloud_to_range_image(pc, origin, True, H, W)
cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
for _ in range(3):
    point_cloud_to_range_image(pc, origin, True, rowMax=H, colMax=W)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    point_cloud_to_range_image(pc, origin, True, rowMax=H, colMax=W)
torch.cuda.synchronize()
e2e_ms = (time.perf_counter() - t0) / 10 * 1e3
# device-resident: call the C ABI directly
lib = cabi.load(); dev = torch.device("cuda:0")
pts = torch.as_tensor(pc).to(dev)
p = cabi.ProjectionParams(); p.n_points, p.point_stride, p.intensity_col, p.height, p.width = n, 4, 3, H, W
p.origin[0], p.origin[1], p.origin[2] = origin
p.dh = math.radians(360) / W; p.dv = math.radians(28) / H; p.h_min = W // (-2) * p.dh + p.dh / 2; p.v_min = math.radians(-25)
f64 = dict(dtype=torch.float64, device=dev)
d, it, ix = torch.empty(H, W, **f64), torch.empty(H, W, **f64), torch.empty(H, W, **f64)
ob, sk = torch.empty(H, W, dtype=torch.uint8, device=dev), torch.empty(H, W, dtype=torch.uint8, device=dev)
ws = torch.empty(int(lib.sdpc_projection_workspace_bytes(H, W)), dtype=torch.uint8, device=dev)
def run():
    lib.sdpc_pointcloud_to_range_image(C.byref(p), C.c_void_p(pts.data_ptr()), C.c_void_p(d.data_ptr()), C.c_void_p(it.data_ptr()),
        C.c_void_p(ob.data_ptr()), C.c_void_p(sk.data_ptr()), C.c_void_p(ix.data_ptr()), C.c_void_p(ws.data_ptr()), ws.numel(),
        C.c_void_p(torch.cuda.current_stream().cuda_stream))
for _ in range(5): run()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50): run()
b.record(); torch.cuda.synchronize()
print(f"N1 {n} points -> {H}x{W}: numpy oracle {cpu_ms:.1f} ms | CUDA via numpy API (h2d+d2h) {e2e_ms:.2f} ms | CUDA device-resident {a.elapsed_time(b)/50*1e3:.1f} us")
