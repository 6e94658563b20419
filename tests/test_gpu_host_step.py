"""GPU: the host-buffer entry point of the C ABI (sdpc_langevin_reproject_step_host, the call bench.py's e2e leg times)
gives bit for bit what the device-pointer calls give: H2D of x, score forward, update, cross-view block, D2H of x and
newImages in one call on one stream."""
import argparse

import numpy as np
import pytest
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200 import cabi
from sdpc_b200.scorenet import NCSN_LiDAR_small
from sdpc_b200.step import StepRunner
from tests.golden import cases

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
N = argparse.Namespace


def _net(H, W, precision):
    cfg = N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
            model=N(ngf=128, num_classes=4, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                    sigma_begin=4.0, sigma_end=0.01, spec_norm=False), device=torch.device(DEV))
    torch.manual_seed(7)
    return NCSN_LiDAR_small(cfg, precision=precision).to(DEV)


@pytest.mark.parametrize("precision", ["bf16", "bf16x3"])
def test_host_step_equals_device_step(precision):
    case = cases.small_multiview("pose")
    B, H, W = case["B"], case["H"], case["W"]
    net = _net(H, W, precision)
    labels = torch.full((B,), 2, device=DEV, dtype=torch.long)
    noise = cases.noise_list(case["x"].shape, 1, 3)[0].to(DEV)

    def runner():
        return StepRunner(case["x"].shape, DEV, case["refer"], case["mask"], case["sky"], case["exist"], case["A"],
                          cabi.SDPC_VARIANT_POSE, to_world=case["toWorld"], from_world=case["fromWorld"])

    # device-pointer path: forward, then sdpc_langevin_reproject_step
    run = runner()
    p = run.params(1e-5, np.sqrt(2e-5), 1.0, case["coef"], 1, True, True, 10.0, False)
    x = case["x"].to(DEV).clone()
    ni = torch.zeros_like(x)
    grad = net(x, labels)
    run.step(p, run.buffers(x, grad, noise, new_images=ni))
    torch.cuda.synchronize()
    # host-buffer path: one call
    run2 = runner()
    x_host = case["x"].clone().pin_memory()
    ni_host = torch.zeros_like(x_host).pin_memory()
    xd, gd, nid = torch.empty_like(x), torch.empty_like(x), torch.zeros_like(x)
    run2.step_host(p, run2.buffers(xd, gd, noise, new_images=nid), x_host, ni_host, scorenet=net, labels=labels)
    torch.cuda.synchronize()
    assert torch.equal(x_host, x.cpu()) and torch.equal(ni_host, ni.cpu())
    assert float(ni_host.abs().max()) > 0
    # pageable host memory and a caller-supplied gradient instead of the score handle
    x_pg = case["x"].clone()
    run3 = runner()
    run3.step_host(p, run3.buffers(xd, gd, noise, new_images=nid), x_pg, None, grad_host=grad.cpu())
    torch.cuda.synchronize()
    assert torch.equal(x_pg, x.cpu())


def test_host_step_rejects_bad_arguments():
    case = cases.small_multiview("pose")
    run = StepRunner(case["x"].shape, DEV, case["refer"], case["mask"], case["sky"], case["exist"], case["A"],
                     cabi.SDPC_VARIANT_POSE, to_world=case["toWorld"], from_world=case["fromWorld"])
    p = run.params(1e-5, 1e-3, 1.0, 0.0, 1, False, False, None, False)
    xd = torch.empty(case["x"].shape, device=DEV)
    with pytest.raises(cabi.SdpcError):                 # no host sample
        run.step_host(p, run.buffers(xd, xd, xd), None)
    net = _net(case["H"], case["W"], "bf16")
    with pytest.raises(cabi.SdpcError):                 # score handle without labels
        run.step_host(p, run.buffers(xd, xd, xd), case["x"].clone(), scorenet=net, labels=None)


def test_step_runner_validates_shapes():
    """broadcastable refer / mask are expanded, anything else is rejected before a device pointer is formed"""
    case = cases.small_multiview("pose")
    kw = dict(to_world=case["toWorld"], from_world=case["fromWorld"])
    args = lambda **o: [o.get("refer", case["refer"]), o.get("mask", case["mask"]), o.get("sky", case["sky"]),
                        o.get("exist", case["exist"])]
    run = StepRunner(case["x"].shape, DEV, *args(mask=case["mask"][:, :1]), case["A"], cabi.SDPC_VARIANT_POSE, **kw)
    assert tuple(run.mask.shape) == tuple(case["x"].shape) and torch.equal(run.mask.cpu(), case["mask"])
    for bad in (dict(refer=case["refer"][:, :, :8]), dict(sky=case["sky"][:2]), dict(exist=case["exist"][:, :8]),
                dict(mask=case["mask"][:3])):
        with pytest.raises(ValueError):
            StepRunner(case["x"].shape, DEV, *args(**bad), case["A"], cabi.SDPC_VARIANT_POSE, **kw)
    with pytest.raises(ValueError):
        StepRunner(case["x"].shape, DEV, *args(), case["A"], cabi.SDPC_VARIANT_POSE, to_world=case["toWorld"][:2],
                   from_world=case["fromWorld"])
