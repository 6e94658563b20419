#!/usr/bin/env python
"""Golden fixture for the two other datasets of row N2 from the reference's OWN source lines:

  * `KITTI360_im_AllForOne`                  LiDARGen/datasets/kitti360_im_AllForOne.py:94-355        (Inpainting.yml)
  * `KITTI360_im_simultaneous_densification` LiDARGen/datasets/kitti360_im_simultenous_densification.py (Densification.yml)

Same method as make_golden_n2.py: the classes cannot be imported (h5py, /data/KITTI-360), so the body of `__getitem__` is
read from /root/reference and executed with a stand-in `self` on the synthetic drive of tests/golden/cases.py, against the
unmodified `point_cloud_to_range_image`.  Run in the build container only:
    python tests/golden/make_golden_n2_variants.py"""
import contextlib
import io
import os
import sys
import types

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import cases  # noqa: E402
from tests.golden.make_golden_n2 import ref_lu, reference_pose_chain  # noqa: E402

DATASETS = "/root/reference/LiDARGen/datasets/"
FILES = {"afo": DATASETS + "kitti360_im_AllForOne.py", "den": DATASETS + "kitti360_im_simultenous_densification.py"}
ITEMS = {"afo": (4, 8), "den": (0, 4)}        # idx = frame * 3 + view: (frame 1, view 1), (frame 2, view 2), (frame 0, view 0)


def reference_item(path, idx, frames, table):
    src = open(path).read().splitlines()
    a = next(i for i, l in enumerate(src) if l.strip().startswith("def __getitem__"))
    b = next(i for i, l in enumerate(src) if l.strip().startswith("def loadVelodyneData"))
    body = "def getitem(self, idx):\n" + "\n".join(src[a + 1:b])
    env = {"np": np, "point_cloud_to_range_image": ref_lu.point_cloud_to_range_image, "print": lambda *a, **k: None}
    exec(body.replace("\n    ", "\n", 0), env)
    H, W = cases.N2_SHAPE
    me = types.SimpleNamespace(full_list=["/synthetic/0000000000.bin"], batchSize=cases.N2_BATCH, frames=frames,
                               Tr_pose_world=table, return_remission=True, rowMax=H, colMax=W, random_roll=False,
                               saveNum=0, modifications=np.array(cases.N2_MODIFICATIONS),
                               loadVelodyneData=lambda name: cases.n2_scan(int(os.path.basename(name)[:-4])))
    with contextlib.redirect_stdout(io.StringIO()):
        return env["getitem"](me, idx)


def main():
    frames, table = reference_pose_chain()
    arrs = {}
    for kind, path in FILES.items():
        for idx in ITEMS[kind]:
            real, known, notsky, index, toW, fromW, goal, toOG, scan_no = reference_item(path, idx, frames, table)
            t = f"{kind}{idx}:"
            arrs.update({t + "real": real, t + "known": np.packbits(known), t + "notsky": np.packbits(notsky),
                         t + "index": np.asarray(index).astype(np.int32), t + "toWorld": toW, t + "fromWorld": fromW,
                         t + "goal": goal, t + "toOGView": toOG, t + "scan": np.int32(scan_no)})
    path = os.path.join(HERE, "dataset_assembly_variants.npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
