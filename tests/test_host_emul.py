"""CPU tests of the kernels' per-element arithmetic (csrc/crossview_core.h) and of the Python
marshalling (StepRunner -> sdpc_step_params / sdpc_step_buffers), through a serial g++ host
emulation of the CUDA kernels (tests/host_emul).  The emulation is test infrastructure only."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200 import cabi
from sdpc_b200.step import StepRunner, translation_origins
from oracle import crossview_ref as cv
from oracle import samplers_ref as sr
from tests.golden import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    out = os.path.join(tempfile.mkdtemp(prefix="sdpc_emul_"), "libemul.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host_emul", "crossview_host.cpp")])
    lib = C.CDLL(out)
    lib.emul_langevin_reproject_step.restype = C.c_int
    lib.emul_langevin_reproject_step.argtypes = [C.POINTER(cabi.StepParams), C.POINTER(cabi.StepBuffers)]
    return lib


class _NoLib:
    pass


def _runner(kind, case):
    kw = {}
    if kind == "pose":
        kw = dict(to_world=case["toWorld"], from_world=case["fromWorld"])
    else:
        kw = dict(origins=translation_origins(case["mods"]))
    return StepRunner(case["x"].shape, "cpu", case["refer"], case["mask"], case["sky"], case["exist"], case["A"],
                      cabi.SDPC_VARIANT_POSE if kind == "pose" else cabi.SDPC_VARIANT_TRANSLATION,
                      lib=_NoLib(), debug=True, **kw)


@pytest.mark.parametrize("kind,sigma,setting", [("pose", 7.5, 5), ("pose", 0.3, 5), ("pose", 0.3, 1),
                                                 ("trans", 7.5, 7), ("trans", 0.3, 4), ("trans", 0.3, 8)])
def test_emulated_kernels_match_oracle(emul, kind, sigma, setting):
    case = cases.small_multiview(kind)
    run = _runner(kind, case)
    sm = sigma if sigma > 1 else 1
    if kind == "pose":
        okw = dict(to_world=case["toWorld"].squeeze(), from_world=case["fromWorld"].squeeze(),
                   min_depth_filter=(setting == 5), controlled_average=True, allowance=10.0)
        p = run.params(0.0, 0.0, 0.0, case["coef"], sm, True, setting == 5, 10.0, False)
    else:
        allow = (5.0 if setting >= 8 else 10.0) if setting >= 7 else None
        okw = dict(origins=cv.translation_origins(case["mods"]), min_depth_filter=True,
                   controlled_average=(setting >= 7), allowance=allow or 10.0, sky_filter=True)
        p = run.params(0.0, 0.0, 0.0, case["coef"], sm, True, True, allow, True)
    geo = cv.make_geometry(case["H"], case["W"])
    ni, im, th, d = cv.shared_images(case["x"], geo, sm, case["A"], case["exist"], case["sky"], return_debug=True, **okw)
    x_ref = cv.apply_correction(case["x"], ni, im, case["sky"], case["mask"], th, case["coef"])

    x = case["x"].clone()
    new_images = torch.zeros_like(x)
    b = run.buffers(x, None, None, new_images=new_images)
    assert emul.emul_langevin_reproject_step(C.byref(p), C.byref(b)) == 0
    dbg = run.debug
    valid = d["valid"]
    # glibc powf vs torch's vectorised powf may differ in the last ulp of the decoded range:
    # allow a handful of boundary flips, everything else bit-exact
    flips = int(((dbg["row"] != d["row"]) | (dbg["col"] != d["col"])).sum())
    assert flips <= 3, flips
    assert int((dbg["valid"].bool() != valid).sum()) <= 3
    assert int((dbg["cnt"] != d["cnt"].int()).sum()) <= 6
    same = dbg["cnt"] == d["cnt"].int()
    assert int(((dbg["winner"] != d["winner"].int()) & same).sum()) <= 3
    assert torch.allclose(dbg["min_d"][same], d["min_d"][same], rtol=1e-6, atol=1e-9)
    bad = (new_images - ni).abs() > 1e-5
    assert int(bad.sum()) <= 8
    assert int(((x - x_ref).abs() > 1e-5).sum()) <= 8
    assert int(run.too_high.item()) == int(bool(th))


def test_emulated_update_is_bit_exact(emul):
    case = cases.small_multiview("pose")
    run = _runner("pose", case)
    sig = cases.short_sigmas()
    noise = cases.noise_list(case["x"].shape, 1, 5)[0]
    grad = cases.fake_score(sig)(case["x"], torch.tensor([1] * case["B"]))
    grad[0, 0, 0, 0] = float("nan")
    grad[0, 0, 0, 1] = float("inf")
    step_size, noise_scale = sr._step_constants(6.2e-6, sig[1], sig[-1])
    ref, gl_ref = sr.langevin_update(case["x"], torch.nan_to_num(grad), case["refer"], case["mask"], noise,
                                     step_size, noise_scale, 1)
    x = case["x"].clone()
    gl = torch.zeros_like(x)
    p = run.params(step_size, noise_scale, 1, 0.0, 1.3, False, False, None, False)
    b = run.buffers(x, grad, noise, grad_likelihood=gl)
    assert emul.emul_langevin_reproject_step(C.byref(p), C.byref(b)) == 0
    assert torch.equal(x, ref) and torch.equal(gl, gl_ref)


def test_emulated_sharded_targets_equal_full(emul):
    """resolving target views in two halves (as two ranks would) equals one full call."""
    case = cases.small_multiview("pose")
    outs = []
    for parts in ([(0, 4)], [(0, 2), (2, 2)], [(0, 1), (1, 3)]):
        x = case["x"].clone()
        ni = torch.zeros_like(x)
        x_in = case["x"].clone()
        for first, count in parts:
            run = _runner("pose", case)
            run.tgt_first, run.tgt_count = first, count
            p = run.params(0.0, 0.0, 0.0, case["coef"], 1, True, True, 10.0, False)
            xx = x_in.clone()
            b = run.buffers(xx, None, None, new_images=ni)
            assert emul.emul_langevin_reproject_step(C.byref(p), C.byref(b)) == 0
            x[first:first + count] = xx[first:first + count]
        outs.append((x, ni.clone()))
    for x, ni in outs[1:]:
        assert torch.equal(x, outs[0][0]) and torch.equal(ni, outs[0][1])


def _emul_share(emul, case, sigma=0.3, setting=5):
    run = _runner("pose", case)
    p = run.params(0.0, 0.0, 0.0, case["coef"], sigma if sigma > 1 else 1, True, setting == 5, 10.0, False)
    x = case["x"].clone()
    ni = torch.full_like(x, 7.0)
    b = run.buffers(x, None, None, new_images=ni)
    assert emul.emul_langevin_reproject_step(C.byref(p), C.byref(b)) == 0
    return x, ni, run


def test_emulated_edge_no_source_pixel_exists(emul):
    """existMask all False: no candidate reaches any z-buffer - the shared images are zero, nothing is corrected"""
    case = cases.small_multiview("pose")
    case["exist"] = torch.zeros_like(case["exist"])
    x, ni, run = _emul_share(emul, case)
    assert int(run.debug["cnt"].abs().sum()) == 0 and int((run.debug["winner"] != -1).sum()) == 0
    assert float(ni.abs().max()) == 0.0
    assert torch.equal(x, case["x"]) and int(run.too_high.item()) == 0


def test_emulated_edge_every_pixel_known(emul):
    """refer_mask all ones: the shared images do not depend on the mask, the correction (1 - mask) vanishes"""
    case = cases.small_multiview("pose")
    _, ni_ref, _ = _emul_share(emul, case)
    case["mask"] = torch.ones_like(case["mask"])
    x, ni, _ = _emul_share(emul, case)
    assert torch.equal(ni, ni_ref) and float(ni.abs().max()) > 0.0
    assert torch.equal(x, case["x"])


def _identity_case(B=2, H=16, W=64, seed=9):
    """B views of one group at the SAME pose: smooth positive ranges, a little per-view noise, every pixel exists"""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = cases.smooth_range_image(1, H, W, seed)
    base[:, 0] = base[:, 0].clamp(0.25, 0.9)                      # away from the min-depth filter and from the far end
    x = (base.repeat(B, 1, 1, 1) + torch.from_numpy(rng.normal(0, 1e-3, size=(B, 2, H, W)).astype(np.float32))).contiguous()
    eye = torch.eye(4, dtype=torch.float64).reshape(1, 1, 4, 4).repeat(B, 1, 1, 1)
    return dict(B=B, A=B, H=H, W=W, R=cases.big_rows(H), x=x, refer=x.clone(), coef=0.25,
                mask=torch.zeros(B, 2, H, W, dtype=torch.int32), sky=torch.ones(B, 1, H, W, dtype=torch.bool),
                exist=torch.ones(B, H, W, dtype=torch.bool), toWorld=eye, fromWorld=eye.clone())


def test_emulated_identity_poses_shift_the_group_mean_down_one_row(emul):
    """the reference's hidden invariant: with identical poses every view receives the mean of the group's
    views shifted DOWN by one row - row 0 stays empty, source row H-1 is dropped - because verticalMin and bigRowMin
    floor half a pixel apart (KITTISampling.py:68-78)"""
    case = _identity_case()
    x, ni, run = _emul_share(emul, case)
    mean = case["x"].double().mean(0, keepdim=True).float()
    assert float(ni[:, :, 0].abs().max()) == 0.0                                  # row 0: no candidate lands there
    assert torch.allclose(ni[:, :, 1:], mean[:, :, :-1].expand_as(ni[:, :, 1:]), rtol=0, atol=2e-6)
    cnt = run.debug["cnt"]                                                        # [B, R, W] grid: rows R-H .. R-1 are the crop
    crop = cnt[:, case["R"] - case["H"]:]
    assert int(crop[:, 0].sum()) == 0 and bool((crop[:, 1:] == case["B"]).all())
    # correction pulls every view towards the shifted mean with weight coef on unknown pixels (mask = 0)
    want = case["x"] + case["coef"] * (-(case["x"] - ni))
    want[:, :, 0] = case["x"][:, :, 0]
    assert torch.allclose(x, want, rtol=0, atol=1e-6)
