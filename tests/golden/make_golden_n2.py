#!/usr/bin/env python
"""Golden fixture for row N2 (dataset assembly) from the reference's OWN source lines.

`KITTI360_im_8batch` (LiDARGen/datasets/kitti360_im_8Batch.py) cannot be imported here: the datasets package needs h5py
and the constructor reads /data/KITTI-360.  This script reads the class's source from /root/reference and executes
  * lines :49-68 (calibration chain) with np.loadtxt answering from synthetic calibration arrays, and
  * the body of `__getitem__` (:94-304) with a stand-in `self` whose `loadVelodyneData` returns synthetic scans,
against the unmodified `point_cloud_to_range_image` of the reference (loaded as a file).  Run in the build container only:
    python tests/golden/make_golden_n2.py"""
import contextlib
import importlib.util
import io
import os
import sys
import textwrap
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import cases  # noqa: E402

REF = "/root/reference/LiDARGen/datasets/kitti360_im_8Batch.py"
spec = importlib.util.spec_from_file_location("ref_lidar_utils", "/root/reference/LiDARGen/datasets/lidar_utils.py")
ref_lu = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref_lu)


def reference_pose_chain():
    src = open(REF).read().splitlines()
    a = next(i for i, l in enumerate(src) if "veloToCam = np.loadtxt" in l)
    b = next(i for i, l in enumerate(src) if "self.Tr_pose_world[frame] = pose" in l)
    body = textwrap.dedent("\n".join(src[a:b + 1]))
    cam_to_velo, cam_to_pose, poses = cases.n2_calibration()

    class FakeNp(types.SimpleNamespace):
        pass
    fake = FakeNp(**{k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
    fake.loadtxt = lambda path, *a, **k: {"calib_cam_to_velo.txt": cam_to_velo, "calib_cam_to_pose.txt": cam_to_pose,
                                          "poses.txt": poses}[os.path.basename(path)]
    me = types.SimpleNamespace()
    exec(body, {"np": fake, "os": os, "self": me})
    return me.frames, me.Tr_pose_world


def reference_item(idx, frames, table):
    src = open(REF).read().splitlines()
    a = next(i for i, l in enumerate(src) if l.strip().startswith("def __getitem__"))
    b = next(i for i, l in enumerate(src) if l.strip().startswith("def loadVelodyneData"))
    body = "def getitem(self, idx):\n" + "\n".join(src[a + 1:b])
    env = {"np": np, "point_cloud_to_range_image": ref_lu.point_cloud_to_range_image, "print": lambda *a, **k: None}
    exec(textwrap.dedent(body) if not body.startswith("def") else body.replace("\n    ", "\n", 0), env)
    H, W = cases.N2_SHAPE
    me = types.SimpleNamespace(full_list=["/synthetic/0000000000.bin"], batchSize=cases.N2_BATCH, frames=frames,
                               Tr_pose_world=table, return_remission=True, rowMax=H, colMax=W, random_roll=False,
                               saveNum=0,
                               loadVelodyneData=lambda name: cases.n2_scan(int(os.path.basename(name)[:-4])))
    with contextlib.redirect_stdout(io.StringIO()):
        return env["getitem"](me, idx)


def main():
    frames, table = reference_pose_chain()
    arrs = {"frames": frames, "poses": np.stack([table[f] for f in frames])}
    for idx in (1, 5, 8):                        # views 1, 2 of frames 0, 1 and view 2 of frame 2
        real, known, notsky, index, toW, fromW, goal, toOG, scan_no = reference_item(idx, frames, table)
        t = f"i{idx}:"
        arrs.update({t + "real": real, t + "known": np.packbits(known), t + "notsky": np.packbits(notsky),
                     t + "index": index.astype(np.int32), t + "toWorld": toW, t + "fromWorld": fromW, t + "goal": goal,
                     t + "toOGView": toOG, t + "scan": np.int32(scan_no)})
    path = os.path.join(HERE, "dataset_assembly.npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
