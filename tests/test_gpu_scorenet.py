"""GPU parity tests of the score network (NCSN_LiDAR_small.forward) through the C ABI.

Oracle: oracle/scorenet_ref.py (torch fp32, pinned bit-for-bit on the reference) and the golden
fixture tests/golden/scorenet_small.npz produced by the unmodified reference module.
Tolerances (max abs error / max abs value of the reference tensor), set at about 1.5x what a B200 measures
(round 2, 16x64 fixture / 64x1024 oracle) so that a regression of the arithmetic is caught:
  fp32  (CUDA-core FMA)            5e-5   measured 2.3e-5 / 2.9e-5
  bf16x3 (tcgen05, hi/lo split)    3e-4   measured 1.5e-4 / 1.7e-4 - the tensor-core arm inside north_star's 1e-3 fp32 bound
  tf32  (tcgen05 kind::tf32)       1.2e-2 measured 6.3e-3 / 7.8e-3 (what cuDNN's default TF32 convs give the reference on a GPU)
  fp16  (tcgen05 kind::f16, half)  1.5e-2 measured 6.9e-3 / 9.9e-3: the same class as tf32, at the bf16 arm's rate
  bf16  (tcgen05 kind::f16, bf16)  8e-2   measured 6.3e-2 / 5.7e-2, stated separately as north_star asks
Per-block intermediates (13 taps) have their own bounds, about 2x the largest measured tap error of the arm.
"""
import argparse
import os

import numpy as np
import pytest
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200.scorenet import NCSN_LiDAR_small
from oracle.scorenet_ref import score_forward
from oracle.weights import make_state_dict
from tests.golden import cases

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
DEV = "cuda:0"
N = argparse.Namespace
NORTH_STAR_FP32 = 1e-3                                                     # score within 1e-3 relative in fp32
TOL = {"fp32": 5e-5, "bf16x3": 3e-4, "tf32": 1.2e-2, "fp16": 1.5e-2, "bf16": 8e-2}         # output
TAP_TOL = {"fp32": 1e-5, "bf16x3": 3e-4, "tf32": 3e-3, "fp16": 3e-3, "bf16": 2e-2}       # per-block taps (largest measured: 4.9e-6, 1.5e-4, 1.6e-3, 1.6e-3, 1.1e-2)
assert TOL["fp32"] <= NORTH_STAR_FP32 and TOL["bf16x3"] <= NORTH_STAR_FP32
TAPS = ["begin_conv", "res1.0", "res1.1", "res2.0", "res2.1", "res3.0", "res3.1", "res4.0", "res4.1",
        "refine1", "refine2", "refine3", "refine4"]


def _cfg(H, W, L=232):
    return N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
             model=N(ngf=128, num_classes=L, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                     sigma_begin=50, sigma_end=0.01, spec_norm=False), device=DEV)


def _net(H, W, precision, keep_taps=False):
    if keep_taps:
        os.environ["SDPC_KEEP_TAPS"] = "1"
    try:
        net = NCSN_LiDAR_small(_cfg(H, W), precision=precision).to(DEV)
        net.load_state_dict(make_state_dict())
        return net
    finally:
        os.environ.pop("SDPC_KEEP_TAPS", None)


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "tf32", "fp16", "bf16"])
def test_small_forward_vs_reference_golden(precision):
    g = np.load(os.path.join(G, "scorenet_small.npz"))
    x, y = torch.from_numpy(g["x"]).to(DEV), torch.from_numpy(g["y"]).to(DEV)
    net = _net(16, 64, precision, keep_taps=True)
    out = net(x, y)
    torch.cuda.synchronize()
    report = []
    for name in TAPS:
        t = cases.subsample_tap(net.read_tap(name, x)).cpu()
        report.append((name, _rel(t, torch.from_numpy(g["tap:" + name]))))
    err = _rel(out.cpu(), torch.from_numpy(g["out"]))
    print(f"[{precision}] per-block rel err:", " ".join(f"{n}={e:.1e}" for n, e in report), f"| out={err:.2e}")
    for name, e in report:
        assert e <= TAP_TOL[precision], (name, e)
    assert err <= TOL[precision], err


@pytest.mark.parametrize("precision", ["fp32", "bf16x3", "tf32", "fp16", "bf16"])
def test_batch_and_reuse_consistency(precision):
    """views are independent (InstanceNorm is per sample): a batched forward equals per-view forwards,
    and buffer reuse across the plan does not leak between runs."""
    net = _net(16, 64, precision)
    x, _ = cases.scorenet_input(16, 64, B=2, seed=5)
    x = torch.cat([x, x.flip(0), x[:1] * 0.5], 0).to(DEV)
    y = torch.tensor([0, 100, 231, 7, 50], device=DEV)
    out = net(x, y)
    out2 = net(x, y)
    assert torch.equal(out, out2)
    for i in range(x.shape[0]):
        single = net(x[i:i + 1], y[i:i + 1])
        assert _rel(single, out[i:i + 1]) <= 1e-6


@pytest.mark.parametrize("precision,H,W", [("fp32", 32, 128), ("bf16x3", 64, 1024), ("tf32", 64, 1024), ("fp16", 64, 1024), ("bf16", 64, 1024)])
def test_larger_forward_vs_oracle(precision, H, W):
    """full-size input (the 128-wide TMA box and multi-tile scheduling) against the fp32 oracle on the device."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = make_state_dict()
    net = _net(H, W, precision)
    x, y = cases.scorenet_input(H, W, B=2, seed=123)
    x, y = x.to(DEV), y.to(DEV)
    out = net(x, y)
    ref = score_forward({k: v.to(DEV) for k, v in sd.items()}, x, y)
    err = _rel(out, ref)
    print(f"[{precision} {H}x{W}] out rel err vs fp32 oracle: {err:.2e}")
    assert err <= TOL[precision], err


def test_state_dict_roundtrip_and_ema():
    from sdpc_b200.ema import EMAHelper
    net = _net(16, 64, "fp32")
    x, y = cases.scorenet_input(16, 64)
    x, y = x.to(DEV), y.to(DEV)
    a = net(x, y)
    dp = torch.nn.DataParallel(net, device_ids=[0])
    states = [dp.state_dict(), None, 0, 0, {k: v.clone() * 1.01 for k, v in net.named_parameters()}]
    dp.load_state_dict(states[0], strict=True)
    ema = EMAHelper(mu=0.999)
    ema.register(dp)
    ema.load_state_dict(states[-1])
    ema.ema(dp)
    b = dp(x, y)
    assert not torch.equal(a, b)                       # the EMA weights reached the kernels
    net.load_state_dict(make_state_dict())
    assert torch.equal(net(x, y), a)


@pytest.mark.parametrize("shape", [(32, 128, 2), (64, 1024, 1)])
def test_conv_kernel_variants_agree(shape, tmp_path):
    """The shipped convolution path (swapped operands, 2-CTA clusters: CTA pairs on one cta_group::2 MMA for Cout = 256,
    TMA multicast of the weights for Cout = 128) against its own fallbacks, each in a fresh process: multicast clusters
    without cta_group::2 and single CTAs (what odd tile counts use) must be bit-identical - same
    tiles, same K order - and the unswapped 128-pixel x 256-channel tiling (statistics summed in another order) stays within the arm's tolerance.  32x128 exercises the two-row half box of the multicast, 64x1024 the half-row one."""
    import subprocess
    import sys
    H, W, B = shape
    probe = os.path.join(os.path.dirname(__file__), "_variant_probe.py")
    outs = {}
    for tag, env in (("default", {}), ("multicast", {"SDPC_CTA2": "0"}), ("single_cta", {"SDPC_CLUSTER": "0"}),
                     ("unswapped", {"SDPC_SWAP256": "0", "SDPC_CLUSTER": "0"})):
        path = str(tmp_path / (tag + ".npy"))
        e = dict(os.environ)
        e.update(env)
        subprocess.run([sys.executable, probe, "bf16", str(H), str(W), str(B), path], check=True, env=e, timeout=600)
        outs[tag] = np.load(path)
    assert np.isfinite(outs["default"]).all()
    assert np.array_equal(outs["default"], outs["multicast"])
    assert np.array_equal(outs["default"], outs["single_cta"])
    scale = np.abs(outs["default"]).max()
    dev = np.abs(outs["default"] - outs["unswapped"]).max() / scale
    print(f"[variants {H}x{W}] swapped vs unswapped tiling: max rel dev {dev:.2e}")
    # both are bf16-arm results: the statistics' summation order moves a few operands by one bf16 ulp, and through 75
    # layers that grows to the arm's own error level (5e-2 against fp32): bound it with the arm's tolerance
    assert dev < TOL["bf16"]
