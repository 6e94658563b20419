"""Helper for tests/test_gpu_scorenet.py::test_conv_kernel_variants_agree: one score-network forward in a fresh process
(the kernel-variant switches SDPC_CLUSTER / SDPC_SWAP256 are read once per process), output saved as .npy.

    python tests/_variant_probe.py <precision> <H> <W> <B> <out.npy>
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200.scorenet import NCSN_LiDAR_small
from oracle.weights import make_state_dict


def main():
    prec, H, W, B, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    NS = argparse.Namespace
    dev = "cuda:0"
    cfg = NS(data=NS(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
             model=NS(ngf=128, num_classes=10, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                      sigma_begin=50, sigma_end=0.01, spec_norm=False), device=dev)
    net = NCSN_LiDAR_small(cfg, precision=prec).to(dev)
    net.load_state_dict(make_state_dict(num_classes=10))
    g = torch.Generator().manual_seed(5)
    x = torch.rand(B, 2, H, W, generator=g).to(dev)
    y = torch.arange(B, device=dev, dtype=torch.long) % 10
    np.save(out, net(x, y).cpu().numpy())


if __name__ == "__main__":
    main()
