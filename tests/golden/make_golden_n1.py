#!/usr/bin/env python
"""Golden fixtures for row N1 (point cloud -> range image) from the UNMODIFIED reference function
/root/reference/LiDARGen/datasets/lidar_utils.py:54 (imported as a file: the datasets package itself needs h5py).
Run in the build container only:  python tests/golden/make_golden_n1.py"""
import contextlib
import importlib.util
import io
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from tests.golden import cases  # noqa: E402

spec = importlib.util.spec_from_file_location("ref_lidar_utils", "/root/reference/LiDARGen/datasets/lidar_utils.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)


def main():
    arrs = {}
    for tag, (n, H, W, seed) in cases.N1_CASES.items():
        pc, origin = cases.synthetic_scan(n, seed)
        with contextlib.redirect_stdout(io.StringIO()):
            d, i, obf, _, sky, idx = ref.point_cloud_to_range_image(pc.copy(), origin.copy(), True, rowMax=H, colMax=W)
        arrs.update({f"{tag}:depth": d, f"{tag}:intensity": i, f"{tag}:obfuscation": obf, f"{tag}:sky": sky,
                     f"{tag}:index": idx.astype(np.int32)})
    path = os.path.join(HERE, "lidar_projection.npz")
    np.savez_compressed(path, **arrs)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
