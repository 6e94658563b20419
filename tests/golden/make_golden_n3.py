"""Golden vectors for row N3 from the reference's own statements.

`visualize_tensor` (LiDARGen/visualization.py:12-62) cannot be imported here (open3d, cv2, matplotlib are absent), so
this script reads the function's source from /root/reference, keeps the numeric statements (:14-43, dropping the
cv2.resize / plt.cm.inferno lines that only prepare the rendering) and executes THEM on a seeded image.  Run in the
build container only; the fixture travels, /root/reference does not.

    python tests/golden/make_golden_n3.py
"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/LiDARGen/visualization.py"


def case_image(seed=7):
    rng = np.random.default_rng(seed)
    img = np.empty((2, 64, 1024), dtype=np.float32)
    img[0] = rng.uniform(-0.05, 1.05, size=(64, 1024)).astype(np.float32)      # depths from below 0.5 m to beyond 63 m
    img[1] = rng.uniform(0.0, 1.0, size=(64, 1024)).astype(np.float32)
    return img


def reference_points(image):
    src = open(REF).read().splitlines()
    start = next(i for i, l in enumerate(src) if l.startswith("def visualize_tensor"))
    stop = next(i for i, l in enumerate(src) if "xyz = pts[mask, :]" in l)
    body = [l[4:] for l in src[start + 1:stop + 1]]
    body = [l for l in body if not re.search(r"cv2\.|plt\.", l)]
    env = {"np": np, "image": image}
    exec("\n".join(body), env)
    return env["xyz"], env["mask"], env["depth"]


def main():
    img = case_image()
    xyz, mask, depth = reference_points(img)
    np.savez_compressed(os.path.join(HERE, "n3_points.npz"), seed=7, count=int(mask.sum()), mask=np.packbits(mask),
                        xyz_every37=xyz[::37], xyz_sum=xyz.sum(0), xyz_abs_sum=np.abs(xyz).sum(0),
                        depth_every101=depth[::101])
    print("points", xyz.shape, "valid", int(mask.sum()))


if __name__ == "__main__":
    main()
