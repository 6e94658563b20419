"""CPU: the runner-layer helpers around the hot path against golden vectors produced by the reference's own
statements / module (tests/golden/make_golden_runner.py): existTotal mask preprocessing, the [2B,3,H,W] layout of the
saved arrays, and the EMA helper that turns a checkpoint's shadow weights into the sampling weights."""
import argparse
import os

import numpy as np
import torch

import sdpc_b200  # noqa: F401
from sdpc_b200 import runner
from sdpc_b200.ema import EMAHelper
from tests.golden.make_golden_runner import exist_counts, perturb, sample_images, small_module

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "runner_helpers.npz"))
NS = argparse.Namespace


def test_exist_mask_preprocessing_matches_reference_statements(tmp_path):
    path = str(tmp_path / "existTotal.npy")
    np.save(path, exist_counts())
    cfg = NS(data=NS(image_size=64, image_width=1024), device="cpu", b200=NS(exist_mask=path))
    m = runner.exist_mask(cfg, 3)
    assert m.dtype == torch.bool and tuple(m.shape) == tuple(G["exist_shape"])
    assert np.array_equal(np.packbits(m.numpy()), G["exist"])
    # the fallback erosion (no scipy) is the same operator
    raw = exist_counts() > np.max(exist_counts()) / 3
    import scipy.ndimage
    want = scipy.ndimage.binary_erosion(raw[2:], border_value=1, iterations=4)
    p = raw[2:].copy()
    for _ in range(4):
        q = np.pad(p, 1, constant_values=True)
        p = q[1:-1, 1:-1] & q[:-2, 1:-1] & q[2:, 1:-1] & q[1:-1, :-2] & q[1:-1, 2:]
    assert np.array_equal(p, want)


def test_saved_array_layout_matches_reference_statements():
    out = runner._to_grid_layout(sample_images())
    assert np.array_equal(out.numpy(), G["grid"])


def test_ema_helper_matches_reference_module():
    m = small_module()
    h = EMAHelper(mu=0.9)
    h.register(torch.nn.DataParallel(m))
    for step in range(3):
        perturb(m, step)
        h.update(m)
    shadow = h.state_dict()
    assert sorted(shadow) == sorted(k[7:] for k in G.files if k.startswith("shadow:"))      # frozen parameter skipped
    for k, v in shadow.items():
        assert np.array_equal(v.numpy(), G["shadow:" + k]), k
    twin = small_module(seed=7)
    h2 = EMAHelper(mu=0.9)
    h2.register(twin)
    h2.load_state_dict(shadow)
    h2.ema(twin)
    for k, v in twin.state_dict().items():
        assert np.array_equal(v.detach().numpy(), G["after:" + k]), k


def test_sensor_geometry_matches_reference_statements():
    """constants and angle tables handed to the kernels against the sampler's own statements (KITTISampling.py:29-102)"""
    from sdpc_b200.geometry import sensor_geometry
    for H, W in ((64, 1024), (16, 64)):
        g = sensor_geometry(H, W, "cpu")
        t = f"geo{H}x{W}:"
        assert g.dh == float(G[t + "horizontalAngles"]) and g.dv == float(G[t + "verticalAngles"])
        assert g.h_min == float(G[t + "horizontalMin"]) and g.v_min == float(G[t + "verticalMin"])
        assert g.big_row_min == float(G[t + "bigRowMin"]) and g.R == int(G[t + "bigRowCount"])
        az, el = torch.from_numpy(G[t + "azimuth"]).reshape(-1), torch.from_numpy(G[t + "elevation"]).reshape(-1)
        assert torch.equal(g.cos_az, torch.cos(az)) and torch.equal(g.sin_az, torch.sin(az))      # KITTISampling.py:176
        assert torch.equal(g.cos_el, torch.cos(el)) and torch.equal(g.sin_el, torch.sin(el))
    assert int(G["geo64x1024:bigRowCount"]) == 114


def test_translation_origins_match_reference_statements():
    """a-5's view origins (models/__init__.py:201-231: an fp32 log2 / pow round trip of |m| divided by m + 1e-8, times 10):
    sign(m) * 10 up to the round trip's last bits, whatever the configured magnitude (SURVEY.md 8a quirk iii)"""
    from sdpc_b200.step import translation_origins
    from tests.golden.make_golden_runner import MODIFICATIONS
    got = translation_origins(torch.tensor(MODIFICATIONS))
    assert got.dtype == torch.float32 and np.array_equal(got.numpy(), G["origins"])
    assert np.allclose(np.abs(G["origins"][np.array(MODIFICATIONS) != 0]), 10.0, atol=1e-4)


def test_exist_mask_on_the_reference_fixture():
    """the only data file the reference ships (MeasureResults/existTotalLiDARGenSettings.npy): after the runner's threshold
    and erosion 68.0 % of the pixels survive and 57 of the 64 rows keep at least one.  Build container only."""
    import pytest
    path = "/root/reference/MeasureResults/existTotalLiDARGenSettings.npy"
    if not os.path.exists(path):
        pytest.skip("the reference tree is not present on this machine")
    cfg = NS(data=NS(image_size=64, image_width=1024), device="cpu", b200=NS(exist_mask=path))
    m = runner.exist_mask(cfg, 2).numpy()
    assert m.shape == (2, 64, 1024) and np.array_equal(m[0], m[1])
    assert abs(m[0].mean() - 0.680) < 5e-4 and int(m[0].any(axis=1).sum()) == 57


def test_runner_output_files_have_the_reference_names_and_layouts(tmp_path):
    """CPU: the files a batch leaves behind (ncsn_runner_kitti_simultaneous.py:650-696,838-893; ncsn_runner_AllForOne.py:
    662-711,905-994): known pixels / ground truth / sky once per batch, the completion (and, for the AllForOne runner, the
    second-to-last entry of the sampler's image list) per ablation arm, each as [2B,3,H,W] float arrays clamped to [0,1] plus
    a PNG grid rendered like `make_grid` + `save_image`"""
    B, Hs, Ws = 4, 8, 32
    cfg = NS(data=NS(channels=2, image_size=Hs, image_width=Ws, logit_transform=False, rescaled=False),
             sampling=NS(batch_size=B, actualBatchSize=2, ckpt_id=897), device="cpu")
    base = runner._Base(NS(image_folder=str(tmp_path), seed=0), cfg)
    g = torch.Generator().manual_seed(5)
    refer = torch.rand(B, 2, Hs, Ws, generator=g, dtype=torch.float64)
    mask = torch.rand(B, 2, Hs, Ws, generator=g) > 0.4
    goal = torch.rand(B, 2, Hs, Ws, generator=g, dtype=torch.float64)
    sky = torch.ones(B, 1, Hs, Ws, dtype=torch.bool)
    base.save_inputs(0, "11_12_", "0", refer, mask, goal, sky)
    load = lambda name: np.load(os.path.join(str(tmp_path), name))
    inp, gt = load("0_11_12__Input_completion_897.pth.npy"), load("0_11_12__GT_completion_897.pth.npy")
    assert inp.shape == (2 * B, 3, Hs, Ws) and gt.shape == (2 * B, 3, Hs, Ws)
    assert np.array_equal(inp[:B, 0], (refer * mask)[:, 0].numpy()) and np.array_equal(inp[B:, 2], (refer * mask)[:, 1].numpy())
    assert np.array_equal(gt[:B, 1], goal[:, 0].numpy())
    assert np.array_equal(load("0_11_12__SKY_897.pth.npy"), sky.numpy())
    outs = [torch.randn(3 * 2 * Hs * Ws, generator=g), torch.randn(3 * 2 * Hs * Ws, generator=g)]       # flat, like the samplers' snapshots viewed by the runner
    masked = base.save_outputs(1, "11_12_", 3, outs, grid_tag="0", nrow=2, shared_initial=True)
    m = load("1_11_12__Masked_completion_897.pth.npy")
    assert m.shape == (6, 3, Hs, Ws) and np.array_equal(m, masked.numpy()) and m.min() >= 0.0 and m.max() <= 1.0
    assert np.array_equal(m[:3, 0], outs[-1].view(3, 2, Hs, Ws)[:, 0].clamp(0, 1).numpy())
    first = load("1_11_12__Shared_completion_initial897.pth.npy")
    assert np.array_equal(first[3:, 0], outs[-2].view(3, 2, Hs, Ws)[:, 1].clamp(0, 1).numpy())
    pngs = sorted(f for f in os.listdir(str(tmp_path)) if f.endswith(".png"))
    assert pngs == ["0_0_GT_image_grid_897.png", "0_0_Input_image_grid_897.png", "1_0_Masked_image_grid_897.png",
                    "1_11_12__Shared_image_grid_initial897.png"]
    from PIL import Image
    im = Image.open(os.path.join(str(tmp_path), "1_0_Masked_image_grid_897.png"))
    assert im.size == (2 * (Ws + 2) + 2, 3 * (Hs + 2) + 2)                       # make_grid: 6 tiles, 2 per row, 2 px padding


def test_get_sigmas_is_bit_equal_to_the_reference_fixture():
    """row a-7: the package's get_sigmas (the runners' and the network buffer's schedule) against sigmas.npz, recorded from
    the unmodified reference's get_sigmas (models/__init__.py:5-18) for the 232- and the 10-level geometric schedules"""
    from sdpc_b200.sigmas import get_sigmas
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "sigmas.npz"))
    for L in (232, 10):
        cfg = NS(model=NS(sigma_dist="geometric", sigma_begin=50, sigma_end=0.01, num_classes=L), device="cpu")
        s = get_sigmas(cfg)
        assert s.dtype == torch.float32 and np.array_equal(s.numpy(), g[f"geometric_{L}"])
    cfg = NS(model=NS(sigma_dist="uniform", sigma_begin=1.0, sigma_end=0.01, num_classes=5), device="cpu")
    assert np.array_equal(get_sigmas(cfg).numpy(), np.linspace(1.0, 0.01, 5).astype(np.float32))
    import pytest
    with pytest.raises(NotImplementedError):
        get_sigmas(NS(model=NS(sigma_dist="other", sigma_begin=1.0, sigma_end=0.01, num_classes=5), device="cpu"))
