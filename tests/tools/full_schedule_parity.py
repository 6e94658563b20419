"""End-to-end parity of a full Line.yml-shape sampling run (V views x 232 noise levels x 5 Langevin steps + denoise):
the B200 sampler (a-4) against the oracle sampler evaluated with torch's own CUDA fp32 kernels (TF32 off) on the same
box, same seed, same Philox noise stream.  Writes a JSON report (profiles/) with per-checkpoint errors and timings."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200 import samplers
from sdpc_b200.scorenet import NCSN_LiDAR_small
from oracle import samplers_ref as sr
from oracle.scorenet_ref import OracleScoreNet
from oracle.sigmas import sigma_schedule
from oracle.weights import make_state_dict

N = argparse.Namespace


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--levels", type=int, default=232)
    ap.add_argument("--steps-each", type=int, default=5)
    ap.add_argument("--arms", default="bf16x3,bf16")
    ap.add_argument("--out", default="gpurun_out/full_schedule_parity.json")
    ap.add_argument("--score", default="hybrid", choices=["hybrid", "raw"],
                    help="raw: the random-init network alone (no restoring force: the chain diverges, trajectories are "
                         "chaotic); hybrid: alpha * network + analytic denoiser -(x - refer)/sigma^2, the stable regime a "
                         "trained score network provides (no checkpoint exists offline)")
    ap.add_argument("--alpha", type=float, default=1e-3)
    a = ap.parse_args()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda:0")
    H, W, V, L = 64, 1024, a.views, a.levels
    from bench import synthetic_group
    g = synthetic_group(V, 1234)
    sig = sigma_schedule(50, 0.01, L).numpy()
    sd = make_state_dict(num_classes=L)
    cfg = N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
            model=N(ngf=128, num_classes=L, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                    sigma_begin=50, sigma_end=0.01, spec_norm=False), device=dev)
    to = lambda t: t.to(dev)
    common = dict(n_steps_each=a.steps_each, step_lr=6.2e-6, existMask=to(g["exist"]), denoise=True, verbose=False,
                  grad_ref=1, correlation_coefficient=0.01)
    sig_t = torch.from_numpy(sig).to(dev)
    refer_d = to(g["refer"])

    def wrap(net):
        if a.score == "raw":
            return net
        return lambda x, y: a.alpha * net(x, y) - (x - refer_d) / (sig_t[y].view(-1, 1, 1, 1) ** 2)
    report = {"score": a.score, "alpha": a.alpha, "views": V, "levels": L, "steps_each": a.steps_each, "sampler": "a-4 (pose matrices), setting 5, minStepToShare 2",
              "oracle": "oracle/samplers_ref.py + scorenet_ref.py with torch CUDA fp32 kernels (allow_tf32=False)", "arms": {}}
    torch.manual_seed(1234)
    t0 = time.time()
    ref_im, _, ref_sh = sr.sampler_pose(to(g["x"]), to(g["refer"]), to(g["mask"]), to(g["sky"]), None, 2, 5, 10,
                                        wrap(OracleScoreNet({k: to(v) for k, v in sd.items()})), sig, g["fromWorld"],
                                        g["toWorld"], V, **common)
    torch.cuda.synchronize()
    report["oracle_seconds"] = time.time() - t0
    for arm in a.arms.split(","):
        net = NCSN_LiDAR_small(cfg, precision=arm).to(dev)
        net.load_state_dict(sd)
        torch.manual_seed(1234)
        t0 = time.time()
        im, _, sh = samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
            to(g["x"]), to(g["refer"]), to(g["mask"]), to(g["sky"]), None, 2, 5, 10, wrap(net), sig, g["fromWorld"],
            g["toWorld"], V, **common)
        torch.cuda.synchronize()
        dt = time.time() - t0

        def err(x, y):
            return {"max_abs": float((x - y).abs().max()), "max_rel": float((x - y).abs().max() / y.abs().max()),
                    "rel_l2": float((x - y).norm() / y.norm())}
        r = {"seconds": dt, "view_steps_per_s": V * (L * a.steps_each + 1) / dt, "final_sample": err(im[-1], ref_im[-1]),
             "final_sample_clamped01": err(im[-1].clamp(0, 1), ref_im[-1].clamp(0, 1)),
             "shared_images": [err(x, y) for x, y in zip(sh, ref_sh)],
             "last_level_shared": [err(x, y) for x, y in zip(im[:-1], ref_im[:-1])]}
        report["arms"][arm] = r
        print(arm, json.dumps(r)[:600], flush=True)
        del net
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    json.dump(report, open(a.out, "w"), indent=1)
    print("oracle seconds", report["oracle_seconds"])


if __name__ == "__main__":
    main()
