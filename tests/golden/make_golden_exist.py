"""Writes simultaneous-diffusion-for-pointclouds_b200/data/exist_mask_lidargen.npz: the beam-existence mask the reference
runners derive from the only data file the reference ships (MeasureResults/existTotalLiDARGenSettings.npy), i.e. that
file after the runner's own threshold + erosion statements (ncsn_runner_kitti_simultaneous.py:527-533, executed here
through runner.exist_mask, which tests/test_runner_helpers.py pins on those statements).  Build container only
(/root/reference is not on the GPU box); the output is 8 KiB bit-packed.

    python tests/golden/make_golden_exist.py
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
SRC = "/root/reference/MeasureResults/existTotalLiDARGenSettings.npy"


def main():
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import runner
    NS = argparse.Namespace
    cfg = NS(data=NS(image_size=64, image_width=1024), device="cpu", b200=NS(exist_mask=SRC))
    m = runner.exist_mask(cfg, 1).numpy()[0]
    out = os.path.join(ROOT, "simultaneous-diffusion-for-pointclouds_b200", "data", "exist_mask_lidargen.npz")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    np.savez_compressed(out, packed=np.packbits(m.reshape(-1)), shape=np.array(m.shape))
    print(out, m.shape, f"{m.mean():.4f} of the pixels exist, {int(m.any(axis=1).sum())} rows keep at least one")


if __name__ == "__main__":
    main()
