"""Ad-hoc device timing of the hot-path pieces (development aid; bench.py is the contract)."""
import argparse
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200 import cabi
from sdpc_b200.scorenet import NCSN_LiDAR_small
from sdpc_b200.step import StepRunner
from tests.golden import cases

N = argparse.Namespace
DEV = "cuda:0"


def timeit(fn, warm=2, it=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / it


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    precs = sys.argv[2].split(",") if len(sys.argv) > 2 else ["tf32", "bf16"]
    H, W = 64, 1024
    cfg = N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
            model=N(ngf=128, num_classes=232, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                    sigma_begin=50, sigma_end=0.01, spec_norm=False), device=DEV)
    torch.manual_seed(1234)          # random-init weights of the module itself (timing only)
    x = torch.rand(B, 2, H, W, device=DEV)
    y = torch.full((B,), 100, device=DEV, dtype=torch.long)
    for prec in precs:
        net = NCSN_LiDAR_small(cfg, precision=prec).to(DEV)
        ms = timeit(lambda: net(x, y), warm=2, it=3 if prec == "fp32" else 10)
        fl = net.flops_per_view(x) * B
        print(f"forward {prec} B={B}: {ms:.3f} ms  -> {B / ms * 1e3:.1f} view-fwd/s, {fl / ms / 1e9:.1f} TFLOP/s, launches={net.launch_count(x)}", flush=True)
        del net
    A = int(sys.argv[3]) if len(sys.argv) > 3 else min(B, 8)          # views per group (B / A groups in the call)
    case = cases.full_multiview(B=B, A=A)
    run = StepRunner(case["x"].shape, DEV, case["refer"], case["mask"], case["sky"], case["exist"], A,
                     cabi.SDPC_VARIANT_POSE, to_world=case["toWorld"], from_world=case["fromWorld"])
    xx = case["x"].to(DEV)
    g = torch.randn_like(xx)
    z = torch.randn_like(xx)
    p = run.params(1e-5, 4e-3, 1.0, 0.01, 1.0, True, True, 10.0, False)
    b = run.buffers(xx, g, z)
    ms = timeit(lambda: run.step(p, b), warm=3, it=20)
    print(f"langevin+crossview step B={B} A={A}: {ms * 1e3:.1f} us", flush=True)


if __name__ == "__main__":
    main()
