"""Per-layer device times of the tensor-core convolutions inside one forward (development aid).

    python tools/conv_layers.py [B] [precision] [steps]

Runs the score network eagerly in profiling mode (CUDA events around every conv launch, sdpc_score_set_profiling) and
prints, per convolution of the plan, the median time over `steps` forwards, its TFLOP/s and its epilogue flags.
"""
import collections
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200.scorenet import NCSN_LiDAR_small

N = collections.namedtuple
DEV = "cuda:0"


def main():
    import argparse
    NS = argparse.Namespace
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    H, W = 64, 1024
    cfg = NS(data=NS(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
             model=NS(ngf=128, num_classes=232, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                      sigma_begin=50, sigma_end=0.01, spec_norm=False), device=DEV)
    dump = tempfile.mktemp(suffix=".csv")
    os.environ["SDPC_PROFILE_DUMP"] = dump
    torch.manual_seed(1234)          # random-init weights of the module itself (timing only)
    net = NCSN_LiDAR_small(cfg, precision=prec).to(DEV)
    x = torch.rand(B, 2, H, W, device=DEV)
    y = torch.full((B,), 100, device=DEV, dtype=torch.long)
    net(x, y)
    net.set_profiling(x, True)
    for _ in range(2):
        net(x, y)
    net.profile_collect(x)
    open(dump, "w").close()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        net(x, y)
    b.record()
    torch.cuda.synchronize()
    ms, fl, n = net.profile_collect(x)
    rows = [l.rstrip("\n").rsplit(",", 2) for l in open(dump)]
    per = len(rows) // steps
    tot = 0.0
    print(f"# B={B} {prec} drop={os.environ.get('SDPC_DEV_EPI_DROP', '0')}: eager forward {a.elapsed_time(b) / steps:.3f} ms, "
          f"convs {ms / steps:.3f} ms, {fl / ms / 1e9:.1f} TFLOP/s, {per} conv launches")
    groups = collections.OrderedDict()
    for i in range(per):
        t = np.median([float(rows[s * per + i][2]) for s in range(steps)]) * 1e3
        f = float(rows[i][1])
        tot += t
        name = rows[i][0]
        key = name.split(" ", 1)[1]
        groups.setdefault(key, []).append(t)
        print(f"{i:3d} {t:8.1f} us {f / t / 1e6:7.1f} TF/s  {name}")
    print("# by shape / epilogue")
    for k, v in sorted(groups.items(), key=lambda kv: -sum(kv[1])):
        print(f"{sum(v):9.1f} us  n={len(v):2d}  mean {np.mean(v):7.1f}  {k}")
    print(f"# sum of medians {tot / 1e3:.3f} ms")
    os.unlink(dump)


if __name__ == "__main__":
    main()
