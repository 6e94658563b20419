"""CPU tests: the oracle (oracle/*.py) against the golden fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  Integer outputs (pixel indices,
counts) and z-buffer winners must be bit-exact; float outputs within 1e-5."""
import os

import numpy as np
import pytest
import torch

from oracle import crossview_ref as cv
from oracle import samplers_ref as sr
from oracle.scorenet_ref import score_forward
from oracle.sigmas import sigma_schedule
from oracle.weights import make_state_dict, parameter_inventory
from tests.golden import cases

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_sigmas_bit_exact():
    g = _load("sigmas.npz")
    for L in (232, 10):
        assert np.array_equal(sigma_schedule(50, 0.01, L).numpy(), g[f"geometric_{L}"])
    s = sigma_schedule(50, 0.01, 232).numpy()
    assert s.dtype == np.float32 and (s > 1).sum() == 107


def test_inventory_matches_reference_state_dict():
    inv = parameter_inventory()
    assert len(inv) == 153                                     # + 'sigmas' buffer = 154 keys
    assert sum(int(np.prod(s)) for _, s in inv) == 29_694_082
    names = [n for n, _ in inv]
    assert names[:2] == ["begin_conv.weight", "begin_conv.bias"]
    assert "res2.0.shortcut.conv.weight" in names and "refine4.output_convs.3_2_conv.weight" in names


def test_scorenet_oracle_matches_reference():
    g = _load("scorenet_small.npz")
    P = make_state_dict()
    taps = {}
    out = score_forward(P, torch.from_numpy(g["x"]), torch.from_numpy(g["y"]), taps)
    ref = g["out"]
    assert np.abs(out.numpy() - ref).max() <= 1e-5 * np.abs(ref).max()
    for k in g.files:
        if k.startswith("tap:"):
            t = cases.subsample_tap(taps[k[4:]]).numpy()
            assert np.abs(t - g[k]).max() <= 1e-5 * np.abs(g[k]).max(), k


def _shared(kind, case, sigma, setting):
    geo = cv.make_geometry(case["H"], case["W"])
    sm = sigma if sigma > 1 else 1
    if kind == "pose":
        kw = dict(to_world=case["toWorld"].squeeze(), from_world=case["fromWorld"].squeeze(),
                  min_depth_filter=(setting == 5), controlled_average=True, allowance=case["allowance"])
    else:
        kw = dict(origins=cv.translation_origins(case["mods"]), min_depth_filter=True,
                  controlled_average=(setting >= 7), allowance=(5.0 if setting >= 8 else 10.0), sky_filter=True)
    return geo, cv.shared_images(case["x"], geo, sm, case["A"], case["exist"], case["sky"], return_debug=True, **kw)


def _check_step(kind, g, pre, case, sigma, setting):
    geo, (ni, im, th, d) = _shared(kind, case, sigma, setting)
    W, R = case["W"], geo.R
    assert np.array_equal((W - 1 - d["col"]).numpy(), g[pre + "colr"])
    assert np.array_equal((R - 1 - d["row"]).numpy(), g[pre + "rowr"])
    assert np.array_equal(d["cnt"].numpy(), g[pre + "cnt"])
    assert np.allclose(d["sum_d"].numpy(), g[pre + "sum_d"], rtol=1e-12, atol=1e-12)
    assert np.allclose(d["sum_i"].numpy(), g[pre + "sum_i"], rtol=1e-5, atol=1e-5)
    if pre + "min_d" in g:
        assert int((d["n_tied"] > 1).sum()) == 0
        assert np.array_equal(d["min_d"].numpy(), g[pre + "min_d"])      # z-buffer winners, bit-exact
        assert np.array_equal(d["min_i"].numpy(), g[pre + "min_i"])
    assert np.allclose(ni.numpy(), g[pre + "new_images"], rtol=1e-6, atol=1e-6)
    x2 = cv.apply_correction(case["x"], ni, im, case["sky"], case["mask"], th, case["coef"])
    assert np.allclose(x2.numpy(), g[pre + "x_final"], rtol=1e-6, atol=1e-6)
    return th


@pytest.mark.parametrize("tag", ["hi", "lo", "lo_nofilter"])
def test_crossview_pose_step(tag):
    g = _load("crossview_pose.npz")
    case = cases.small_multiview("pose")
    _check_step("pose", g, tag + ":", case, float(g[tag + ":sigma"]), int(g[tag + ":setting"]))


@pytest.mark.parametrize("tag", ["hi7", "lo7", "lo4", "lo8"])
def test_crossview_translation_step(tag):
    g = _load("crossview_trans.npz")
    case = cases.small_multiview("trans")
    _check_step("trans", g, tag + ":", case, float(g[tag + ":sigma"]), int(g[tag + ":setting"]))


def test_crossview_too_high_gate():
    g = _load("crossview_toohigh.npz")
    case = cases.small_multiview("pose", outlier=True)
    th = _check_step("pose", g, "", case, 0.3, 5)
    assert bool(th)
    assert np.array_equal(g["x_final"], case["x"].numpy())      # correction fully suppressed


def test_translation_origins_collapse_to_pm10():
    o = cv.translation_origins(cases.small_multiview("trans")["mods"])[:, :, 0, 0]
    assert o.dtype == torch.float32
    assert set(np.round(o.numpy().ravel(), 4).tolist()) <= {0.0, 10.0, -10.0}


def test_self_reprojection_shifts_one_row():
    """identical views, identity poses -> shared image = input rolled down one row."""
    H, W, A = 16, 64, 2
    geo = cv.make_geometry(H, W)
    x = cases.smooth_range_image(A, H, W, 5)
    x[1] = x[0]
    eye = torch.eye(4, dtype=torch.float64).repeat(A, 1, 1)
    exist = torch.ones(A, H, W, dtype=torch.bool)
    exist[1:] = False
    sky = torch.ones(A, 1, H, W, dtype=torch.bool)
    ni, im, _ = cv.shared_images(x, geo, 1, A, exist, sky, to_world=eye, from_world=eye,
                                 min_depth_filter=False, controlled_average=True, allowance=10)
    assert not im[0, 0].any()
    assert torch.allclose(ni[0, :, 1:], x[0, :, :-1], atol=1e-6)


def _noise_iter(lst):
    lst = list(lst)
    return lambda x: lst.pop(0)


def test_sampler_pose_trajectory():
    g = _load("sampler_pose.npz")
    sig = cases.short_sigmas()
    case = cases.small_multiview("pose")
    im, tg, sh = sr.sampler_pose(case["x"].clone(), case["refer"], case["mask"], case["sky"], None, 1, 5,
                                 case["allowance"], cases.fake_score(sig), sig, case["fromWorld"], case["toWorld"],
                                 case["A"], n_steps_each=2, step_lr=6.2e-6, existMask=case["exist"], denoise=True,
                                 verbose=False, grad_ref=1, correlation_coefficient=0.01,
                                 noise_fn=_noise_iter(cases.noise_list(case["x"].shape, 8, 77)))
    assert len(im) == int(g["n_images"]) and len(sh) == int(g["n_shared"]) and tg == []
    for i, t in enumerate(im):
        assert np.allclose(t.numpy(), g[f"images{i}"], rtol=1e-5, atol=1e-5)


def test_sampler_translation_trajectory():
    g = _load("sampler_trans.npz")
    sig = cases.short_sigmas()
    case = cases.small_multiview("trans")
    im, tg, sh = sr.sampler_translation(case["x"].clone(), case["refer"], case["mask"], case["sky"], None, 1, 7,
                                        cases.fake_score(sig), sig, case["mods"], case["A"], n_steps_each=2,
                                        step_lr=6.2e-6, existMask=case["exist"], denoise=True, verbose=False,
                                        grad_ref=1, correlation_coefficient=0.01,
                                        noise_fn=_noise_iter(cases.noise_list(case["x"].shape, 8, 78)))
    assert len(im) == int(g["n_images"])
    for i, t in enumerate(im):
        assert np.allclose(t.numpy(), g[f"images{i}"], rtol=1e-5, atol=1e-5)


def test_sampler_single_view_trajectory():
    g = _load("sampler_single.npz")
    sig = cases.short_sigmas()
    case = cases.small_multiview("trans")
    im, tg = sr.sampler_single_view(case["x"].clone(), case["refer"], case["mask"], cases.fake_score(sig), sig,
                                    n_steps_each=2, step_lr=6.2e-6, denoise=True, verbose=False, grad_ref=1,
                                    noise_fn=_noise_iter(cases.noise_list(case["x"].shape, 8, 79)))
    assert len(im) == int(g["n_images"]) == 10 and len(tg) == 1
    for i, t in enumerate(im):
        assert np.array_equal(t.numpy(), g[f"images{i}"])


def test_n4_unconditional_and_densification_trajectories():
    """row N4: the oracle restatements against trajectories of the unmodified reference samplers (models/__init__.py:20-109)."""
    g = _load("sampler_n4.npz")
    sig = cases.short_sigmas()
    case = cases.small_multiview("trans")
    im = sr.sampler_unconditional(case["x"].clone(), cases.fake_score(sig), sig, n_steps_each=2, step_lr=6.2e-6,
                                  final_only=False, denoise=True,
                                  noise_fn=_noise_iter(cases.noise_list(case["x"].shape, 8, 80)))
    assert len(im) == int(g["u_n"]) == 9
    for i, t in enumerate(im):
        assert np.array_equal(t.numpy(), g[f"u{i}"])
    im, tg = sr.sampler_densification(case["x"].clone(), case["refer"], cases.fake_score(sig), sig, n_steps_each=2,
                                      step_lr=6.2e-6, denoise=True, grad_ref=0.1, sampling_step=4,
                                      noise_fn=_noise_iter(cases.noise_list(case["x"].shape, 8, 81)))
    assert len(im) == int(g["d_n"]) == 10 and len(tg) == 1
    for i, t in enumerate(im):
        assert np.array_equal(t.numpy(), g[f"d{i}"])


def test_crossview_full_size_checksums():
    g = _load("crossview_full.npz")
    case = cases.full_multiview()
    geo = cv.make_geometry(64, 1024)
    assert geo.R == 114
    ni, im, th, d = cv.shared_images(case["x"], geo, 1, case["A"], case["exist"], case["sky"],
                                     to_world=case["toWorld"].squeeze(), from_world=case["fromWorld"].squeeze(),
                                     min_depth_filter=True, controlled_average=True, allowance=10, return_debug=True)
    s = cases.FULL_STRIDE
    colr = (1023 - d["col"]).numpy().astype(np.int64)
    rowr = (geo.R - 1 - d["row"]).numpy().astype(np.int64)
    w = np.arange(colr.size) % 1009
    assert int(colr.sum()) == int(g["colr_sum"]) and int(rowr.sum()) == int(g["rowr_sum"])
    assert int((colr.reshape(-1) * w).sum()) == int(g["colr_wsum"])
    assert int((rowr.reshape(-1) * w).sum()) == int(g["rowr_wsum"])
    assert int(d["cnt"].sum()) == int(g["cnt_sum"]) and int((d["cnt"] > 0).sum()) == int(g["n_filled"])
    assert np.array_equal(d["min_d"].numpy().reshape(-1)[::s], g["min_d_s"])
    assert np.allclose(ni.numpy().reshape(-1)[::s], g["new_images_s"], atol=1e-6)


@pytest.mark.parametrize("tag", sorted(cases.FULL_TRANS_RUNS))
def test_crossview_full_size_translation_checksums(tag):
    """a-5 in the shape of BASELINE configs 3 / 4 (V = A = 8 at 64x1024, configured offsets, the shipped existTotal mask,
    inpainting / rows-0::4 densification masks, settings 7 and 8): the oracle against one step of the unmodified reference
    (models/__init__.py:112-602; fixture crossview_full_trans.npz) - every index, every count, the samples of the values."""
    g = _load("crossview_full_trans.npz")
    sigma, setting, densify = cases.FULL_TRANS_RUNS[tag]
    case = cases.full_translation(densify=densify)
    geo = cv.make_geometry(64, 1024)
    sm = sigma if sigma > 1 else 1
    ni, im, th, d = cv.shared_images(case["x"], geo, sm, case["A"], case["exist"], case["sky"],
                                     origins=cv.translation_origins(case["mods"]), min_depth_filter=True,
                                     controlled_average=True, allowance=(5.0 if setting >= 8 else 10.0), sky_filter=True,
                                     return_debug=True)
    x_final = cv.apply_correction(case["x"], ni, im, case["sky"], case["mask"], th, case["coef"])
    s = cases.FULL_STRIDE
    k = lambda n: g[f"{tag}:{n}"]
    colr = (1023 - d["col"]).numpy().astype(np.int64)
    rowr = (geo.R - 1 - d["row"]).numpy().astype(np.int64)
    w = np.arange(colr.size) % 1009
    assert int(colr.sum()) == int(k("colr_sum")) and int(rowr.sum()) == int(k("rowr_sum"))
    assert int((colr.reshape(-1) * w).sum()) == int(k("colr_wsum")) and int((rowr.reshape(-1) * w).sum()) == int(k("rowr_wsum"))
    assert int(d["cnt"].sum()) == int(k("cnt_sum")) and int((d["cnt"] > 0).sum()) == int(k("n_filled"))
    assert np.array_equal(d["cnt"].numpy().reshape(-1)[::s], k("cnt_s"))
    assert np.array_equal(d["min_d"].numpy().reshape(-1)[::s], k("min_d_s"))
    assert np.allclose(ni.numpy().reshape(-1)[::s], k("new_images_s"), atol=1e-6)
    assert np.allclose(x_final.numpy().reshape(-1)[::s], k("x_final_s"), atol=1e-6)
    assert abs(float(np.abs(ni.numpy().astype(np.float64)).sum()) - float(k("new_images_abs_sum"))) < 1e-6 * float(k("new_images_abs_sum"))
