"""Row N2 timing: CUDA assemble_view (host mirror, raw scans in, finished images out) next to the numpy oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
import sdpc_b200  # noqa
from sdpc_b200 import datasets
from oracle import dataset_assembly_ref as da
from tests.golden import cases

cam_to_velo, cam_to_pose, poses = cases.n2_calibration()
frames, table = datasets.velo_to_world_poses(cam_to_velo, cam_to_pose[0], poses)
scan, goal = cases.n2_scan(frames[0], 120000), cases.n2_scan(frames[10], 120000)
t0 = time.perf_counter()
for _ in range(3):
    da.assemble_view(scan, goal, table[frames[0]], table[frames[10]], 64, 1024)
cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
for _ in range(3):
    datasets.assemble_view(scan, goal, table[frames[0]], table[frames[10]])
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    datasets.assemble_view(scan, goal, table[frames[0]], table[frames[10]])
torch.cuda.synchronize()
gpu_ms = (time.perf_counter() - t0) / 10 * 1e3
print(f"N2 one item (2 x 120 000 points -> 64x1024 inputs + ground truth): numpy oracle {cpu_ms:.1f} ms | CUDA via the host mirror "
      f"(h2d of the raw scans, d2h of the images) {gpu_ms:.2f} ms")
