"""Golden vectors for the runner-layer helpers around the hot path, from the reference's own code.

* existTotal mask preprocessing and the saved-array layout: the statements of
  LiDARGen/runners/ncsn_runner_kitti_simultaneous.py (:528-532 and the `maskedSample` lines at :866-868) are read from
  /root/reference at generation time and executed on seeded inputs (`import runners` itself needs h5py / open3d).
* sensor-model constants (angles per pixel, grid minima, bigRowCount) and the azimuth / elevation tables: the statements of
  LiDARGen/models/KITTISampling.py:29-102 executed on stand-in samples.
* EMAHelper: LiDARGen/models/ema.py is imported as a file (it only needs torch) and driven through
  register / update / state_dict / load_state_dict / ema on a small module.
Run in the build container only; the fixture travels, /root/reference does not.

    python tests/golden/make_golden_runner.py
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
RUNNER = "/root/reference/LiDARGen/runners/ncsn_runner_kitti_simultaneous.py"
EMA = "/root/reference/LiDARGen/models/ema.py"
SAMPLER = "/root/reference/LiDARGen/models/KITTISampling.py"
MODELS_INIT = "/root/reference/LiDARGen/models/__init__.py"
# data.modifications of the Inpainting / Densification configurations plus zero, +-3 and a large offset
MODIFICATIONS = [[0, 0, 0], [5, -5, 0], [-5, -5, 0], [0, 5, 0], [-10, 10, 0], [10, 10, 0], [-10, 0, 0], [3, -3, 1], [200, 0, -7]]


def exist_counts(seed=3, H=64, W=1024):
    """per-pixel hit counts like existTotalLiDARGenSettings.npy: smooth field with weak rows and scattered holes"""
    rng = np.random.default_rng(seed)
    v = rng.uniform(0.3, 1.0, size=(H, W)) * 8601.0
    v[:3] *= rng.uniform(0.1, 0.9, size=(3, W))
    v *= (rng.uniform(size=(H, W)) > 0.03)
    return v


def sample_images(seed=4, B=6, H=8, W=16):
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.uniform(-0.2, 1.2, size=(B, 2, H, W)).astype(np.float32))


def small_module(seed=6):
    torch.manual_seed(seed)
    m = torch.nn.Sequential(torch.nn.Conv2d(2, 4, 3), torch.nn.Conv2d(4, 2, 1))
    m[1].bias.requires_grad_(False)                                   # frozen parameters are skipped by the helper
    return m


def perturb(module, step):
    g = torch.Generator().manual_seed(100 + step)
    with torch.no_grad():
        for p in module.parameters():
            p.add_(torch.randn(p.shape, generator=g) * 0.1)


def _lines(path, first_marker, count):
    src = open(path).read().splitlines()
    i = next(k for k, l in enumerate(src) if first_marker in l)
    return [l.strip() for l in src[i:i + count]]


def main():
    import scipy.ndimage
    # ---- existTotal preprocessing: the four statements after np.load (:528-532 minus the comments)
    body = [l for l in _lines(RUNNER, "existVals = existVals > np.max(existVals) / 3", 5) if not l.startswith("#")]
    assert len(body) == 3 and "binary_erosion" in body[1] and "np.tile" in body[2], body
    cfg = type("C", (), {})()
    cfg.config = type("C", (), {})()
    cfg.config.sampling = type("C", (), {"batch_size": 3})()
    env = {"np": np, "scipy": scipy, "existVals": exist_counts(), "self": cfg}
    exec("\n".join(body), env)
    exist = env["existVals"]
    # ---- saved-array layout of a two-channel batch
    body = _lines(RUNNER, "maskedSample = maskedSample.transpose(1, 0)", 3)
    assert "reshape" in body[1] and "torch.cat" in body[2], body
    env = {"torch": torch, "maskedSample": sample_images()}
    exec("\n".join(body), env)
    grid = env["maskedSample"].numpy()
    # ---- EMA helper
    spec = importlib.util.spec_from_file_location("ref_ema", EMA)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    m = small_module()
    h = ref.EMAHelper(mu=0.9)
    h.register(torch.nn.DataParallel(m))
    for step in range(3):
        perturb(m, step)
        h.update(m)
    shadow = {k: v.clone().numpy() for k, v in h.state_dict().items()}
    twin = small_module(seed=7)
    h2 = ref.EMAHelper(mu=0.9)
    h2.register(twin)
    h2.load_state_dict(h.state_dict())
    h2.ema(twin)
    after = {k: v.detach().clone().numpy() for k, v in twin.state_dict().items()}
    # ---- sensor-model constants and angle tables of the samplers: the statements of KITTISampling.py between
    #      `rowMax = x_mod.shape[-2]` and the `elevation = ...` line, executed on stand-in samples of two sizes
    import math
    src = open(SAMPLER).read().splitlines()
    i0 = next(k for k, l in enumerate(src) if l.strip().startswith("rowMax = x_mod.shape[-2]"))
    i1 = next(k for k, l in enumerate(src) if l.strip().startswith("elevation = torch.from_numpy"))
    body = "\n".join(l[4:] if l.startswith("    ") else l for l in src[i0:i1 + 1])
    geo = {}
    for Hh, Ww in ((64, 1024), (16, 64)):
        env = {"np": np, "torch": torch, "math": math, "x_mod": torch.zeros(2, 2, Hh, Ww)}
        exec(body, env)
        for k in ("horizontalAngles", "verticalAngles", "horizontalMin", "verticalMin", "bigRowMin"):
            geo[f"geo{Hh}x{Ww}:{k}"] = np.float64(env[k])
        geo[f"geo{Hh}x{Ww}:bigRowCount"] = np.int64(env["bigRowCount"])
        geo[f"geo{Hh}x{Ww}:azimuth"] = env["azimuth"].numpy()
        geo[f"geo{Hh}x{Ww}:elevation"] = env["elevation"].numpy()
    # ---- a-5's view origins: the statements of models/__init__.py (:201, :215, :224-225, :231) on the configured offsets
    src = open(MODELS_INIT).read().splitlines()
    pick = []
    for marker in ("originListOG = torch.unsqueeze(torch.unsqueeze(modificationList,-1),-1)", "heWhoModsTheModMan = 1",
                   "originList = ((torch.log2(torch.abs(originListOG)+1)) / 6) * heWhoModsTheModMan",
                   "originList = (torch.pow(2,(originList*6))-1)", "originList = originList/(originListOG+0.00000001) * 10"):
        pick.append(next(l.strip() for l in src if l.strip() == marker))
    env = {"torch": torch, "modificationList": torch.tensor(MODIFICATIONS)}
    exec("\n".join(pick), env)
    geo["origins"] = env["originList"].numpy()
    np.savez_compressed(os.path.join(HERE, "runner_helpers.npz"), exist=np.packbits(exist), exist_shape=np.array(exist.shape), **geo,
                        grid=grid, **{"shadow:" + k: v for k, v in shadow.items()},
                        **{"after:" + k: v for k, v in after.items()})
    print("exist", exist.shape, int(exist.sum()), "grid", grid.shape, "shadow keys", sorted(shadow))


if __name__ == "__main__":
    main()
