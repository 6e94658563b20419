"""Strong scaling of ONE group of views sharded over the ranks (BASELINE configs 3/4 layout: the views of a group live on
different GPUs, so every step all-gathers the updated planes and all-reduces the tooHigh maximum).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/time_sharded_sampler.py [views] [levels]

Runs sampler a-4 through the public API (`shard=ViewShard`) on a short schedule and prints view-steps/s of the whole call
(wall clock around the call on rank 0 after a warm-up call, barrier + synchronize on both sides)."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import sdpc_b200  # noqa: F401
from sdpc_b200 import samplers
from sdpc_b200.dist import ViewShard
from sdpc_b200.scorenet import NCSN_LiDAR_small
from sdpc_b200.sigmas import get_sigmas
from tests.golden import cases


def main():
    V = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    NS = argparse.Namespace
    H, W = 64, 1024
    cfg = NS(data=NS(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
             model=NS(ngf=128, num_classes=L, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                      sigma_begin=50, sigma_end=0.01, spec_norm=False), device=dev)
    torch.manual_seed(1234)
    net = NCSN_LiDAR_small(cfg, precision="bf16").to(dev)
    sig = get_sigmas(cfg).cpu().numpy()
    case = cases.full_multiview(B=V, A=V)
    to = lambda t: t.to(dev)
    shard = ViewShard(V, V) if world > 1 else None
    kw = dict(n_steps_each=5, step_lr=6.2e-6, existMask=to(case["exist"]), denoise=True, verbose=False, grad_ref=1,
              correlation_coefficient=0.01)
    if shard is not None:
        kw["shard"] = shard

    def run():
        torch.manual_seed(7)
        return samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
            to(case["x"]), to(case["refer"]), to(case["mask"]), to(case["sky"]), None, 2, 5, 10, net, sig,
            case["fromWorld"], case["toWorld"], V, **kw)

    run()                                    # warm-up: plans, graphs, NCCL communicators
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        fw = L * 5 + 1
        print(f"sharded sampler: {V} views of one group over {world} GPU(s) ({V // world} per rank), {L} levels x 5 steps + denoise = "
              f"{fw} forwards: {dt * 1e3:.1f} ms -> {V * fw / dt:.1f} view-steps/s, {dt / fw * 1e3:.2f} ms per step", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
