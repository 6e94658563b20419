"""A/B aid for kernel-variant switches that are read once per process (SDPC_CTA2, SDPC_CLUSTER, ...): one process per
setting saves score-network outputs on fixed inputs and prints the forward time; `compare` checks two runs bit for bit.

    SDPC_X3_CLUSTER=1 python tools/ab_probe.py run gpurun_out/ab_x3cl
    python tools/ab_probe.py run gpurun_out/ab_def
    python tools/ab_probe.py compare gpurun_out/ab_def gpurun_out/ab_x3cl
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

CASES = [("bf16", 32, 128, 2), ("bf16", 64, 1024, 1), ("tf32", 32, 128, 2), ("bf16x3", 32, 128, 2), ("bf16x3", 64, 1024, 1)]


def run(prefix, time_b):
    import torch
    import sdpc_b200  # noqa: F401
    from sdpc_b200.scorenet import NCSN_LiDAR_small
    NS = argparse.Namespace
    dev = "cuda:0"
    for prec, H, W, B in CASES:
        cfg = NS(data=NS(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
                 model=NS(ngf=128, num_classes=10, nonlinearity="elu", normalization="InstanceNorm++",
                          sigma_dist="geometric", sigma_begin=50, sigma_end=0.01, spec_norm=False), device=dev)
        torch.manual_seed(1234)      # the module's own random init, identical in every process
        net = NCSN_LiDAR_small(cfg, precision=prec).to(dev)
        g = torch.Generator().manual_seed(5)
        x = torch.rand(B, 2, H, W, generator=g).to(dev)
        y = torch.arange(B, device=dev, dtype=torch.long) % 10
        out = net(x, y)
        torch.cuda.synchronize()
        np.save(f"{prefix}_{prec}_{H}x{W}.npy", out.cpu().numpy())
        print(f"[{prefix}] {prec} {H}x{W} B={B}: finite={bool(torch.isfinite(out).all())} absmax={out.abs().max().item():.4g}", flush=True)
        if (prec, H, W) == ("bf16", 64, 1024) and time_b:
            xb = torch.rand(time_b, 2, H, W, device=dev)
            yb = torch.full((time_b,), 5, device=dev, dtype=torch.long)
            for _ in range(3):
                net(xb, yb)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                net(xb, yb)
            b.record()
            torch.cuda.synchronize()
            ms = a.elapsed_time(b) / 20
            print(f"[{prefix}] forward bf16 B={time_b}: {ms:.3f} ms -> {time_b / ms * 1e3:.1f} view-fwd/s", flush=True)
        del net


def compare(pa, pb):
    ok = True
    for prec, H, W, _ in CASES:
        a, b = np.load(f"{pa}_{prec}_{H}x{W}.npy"), np.load(f"{pb}_{prec}_{H}x{W}.npy")
        same = np.array_equal(a, b)
        dev = float(np.abs(a - b).max() / np.abs(a).max())
        print(f"compare {prec} {H}x{W}: identical={same} max rel dev={dev:.3e}", flush=True)
        ok = ok and same
    return ok


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 8)
    else:
        sys.exit(0 if compare(sys.argv[2], sys.argv[3]) else 1)
