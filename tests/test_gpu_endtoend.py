"""GPU: sampler trajectories with the REAL score network against the oracle samplers (torch fp32 kernels on the same
device, TF32 off), same seed -> same Philox noise stream.  north_star: sample values within 1e-3 relative in fp32."""
import argparse

import numpy as np
import pytest
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200 import samplers
from sdpc_b200.scorenet import NCSN_LiDAR_small
from oracle import samplers_ref as sr
from oracle.scorenet_ref import OracleScoreNet
from oracle.sigmas import sigma_schedule
from oracle.weights import make_state_dict
from tests.golden import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N = argparse.Namespace


def _setup(H, W, L):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
            model=N(ngf=128, num_classes=L, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                    sigma_begin=50, sigma_end=0.01, spec_norm=False), device=DEV)
    sd = make_state_dict(num_classes=L)
    return cfg, sd, sigma_schedule(50, 0.01, L).numpy()


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-5), ("bf16x3", 1e-4)])      # measured 3.7e-6 / 2.3e-5 (north_star: 1e-3)
def test_config1_single_view_baseline(precision, tol):
    """BASELINE.json configs[0]: a-6, B=1, 2x64x1024, 10-level geometric schedule (model.num_classes=10), denoise."""
    H, W, L = 64, 1024, 10
    cfg, sd, sig = _setup(H, W, L)
    net = NCSN_LiDAR_small(cfg, precision=precision).to(DEV)
    net.load_state_dict(sd)
    g = cases.smooth_range_image(1, H, W, 3).to(DEV)
    mask = (torch.rand(1, 1, H, W, generator=torch.Generator().manual_seed(5)) < 0.6).int().repeat(1, 2, 1, 1).to(DEV)
    x0 = torch.rand(1, 2, H, W, generator=torch.Generator().manual_seed(6)).to(DEV)
    torch.manual_seed(99)
    im, tg = samplers.anneal_Langevin_dynamics_inpainting(x0, g, mask, net, sig, n_steps_each=2, step_lr=6.2e-6,
                                                          denoise=True, verbose=False, grad_ref=1)
    torch.manual_seed(99)
    ref, _ = sr.sampler_single_view(x0.clone(), g, mask, OracleScoreNet({k: v.to(DEV) for k, v in sd.items()}), sig,
                                    n_steps_each=2, step_lr=6.2e-6, denoise=True, verbose=False, grad_ref=1)
    assert len(im) == len(ref) == 2 * L + 2
    errs = [_rel(a, b) for a, b in zip(im, ref)]
    print(f"[config1 {precision}] max rel err over {len(im)} snapshots: {max(errs):.2e}, final {errs[-1]:.2e}")
    assert max(errs) <= tol


# bounds at about 2x what a B200 measures (round 2): final sample 2.5e-5 (bf16x3) / 8.9e-3 (bf16) of the oracle's, shared-image
# pixels that moved to a neighbouring cell 0.17 % / 1.75 %; fp16 arm: 1.05e-3 and 1.2 %
@pytest.mark.parametrize("precision,tol,max_flip_frac", [("bf16x3", 1e-4, 0.004), ("fp16", 2.5e-3, 0.025), ("bf16", 2e-2, 0.035)])
def test_line_sampler_with_real_network(precision, tol, max_flip_frac):
    """a-4 with the real score network on 4 views (A=4), levels spread over the schedule so that both sigmaMod
    branches and the minStepToShare switch are crossed."""
    H, W, L = 64, 1024, 12
    cfg, sd, sig = _setup(H, W, L)
    net = NCSN_LiDAR_small(cfg, precision=precision).to(DEV)
    net.load_state_dict(sd)
    case = cases.full_multiview(B=4, A=4)
    to = lambda t: t.to(DEV)
    kw = dict(n_steps_each=2, step_lr=6.2e-6, existMask=to(case["exist"]), denoise=True, verbose=False, grad_ref=1,
              correlation_coefficient=0.01)
    torch.manual_seed(7)
    im, _, sh = samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
        to(case["x"]), to(case["refer"]), to(case["mask"]), to(case["sky"]), None, 2, 5, 10, net, sig, case["fromWorld"],
        case["toWorld"], 4, **kw)
    torch.manual_seed(7)
    ref, _, rsh = sr.sampler_pose(to(case["x"]), to(case["refer"]), to(case["mask"]), to(case["sky"]), None, 2, 5, 10,
                                  OracleScoreNet({k: v.to(DEV) for k, v in sd.items()}), sig, case["fromWorld"],
                                  case["toWorld"], 4, **kw)
    assert len(im) == len(ref)
    e_final = _rel(im[-1], ref[-1])
    # The shared image is a function of ROUNDED pixel indices: a sample that differs by 1e-5 moves a re-projected
    # point by ~1e-2 pixel, so a fraction of a percent of the candidates legitimately lands in the neighbouring pixel
    # (the reference shows the same sensitivity between its CPU and GPU runs).  Bit-exactness of the step itself for
    # identical inputs is asserted in test_gpu_crossview.py; here only the flipped fraction is bounded.
    bad = int(((im[0] - ref[0]).abs() > 1e-2 * ref[0].abs().max()).sum())
    print(f"[line {precision}] final sample rel err {e_final:.2e}; shared-image pixels off by >1e-2: {bad} of {ref[0].numel()}")
    assert e_final <= tol
    assert bad <= max_flip_frac * ref[0].numel()
