#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference, read-only):

    python tests/golden/make_golden.py

The reference ships no tests or golden vectors, so these fixtures are what
pins the oracle (oracle/*.py) and, through it, the CUDA path.  The reference
functions are imported from /root/reference/LiDARGen/models and called as-is;
intermediates that the reference does not return (pixel indices, per-pixel
counts, z-buffer winners) are captured by temporarily wrapping the torch entry
points it calls (torch.round, torch.sparse_coo_tensor, torch.randn_like) - the
reference source is never edited or copied.

Fixtures written to tests/golden/:
  sigmas.npz            get_sigmas for (50,0.01,232) and (50,0.01,10)
  scorenet_small.npz    NCSN_LiDAR_small forward, ngf=128, 16x64 input, deterministic weights
  crossview_pose.npz    one cross-view step of a-4 (B=4,A=2,16x64), sigma>1 and sigma<=1, with internals
  crossview_trans.npz   same for a-5 (setting 7 and setting 4)
  crossview_toohigh.npz a-4 step with the tooHigh gate tripped
  sampler_pose.npz      a-4 short schedule trajectory (4 levels x 2 steps, denoise)
  sampler_trans.npz     a-5 short schedule trajectory
  sampler_single.npz    a-6 short schedule trajectory
  crossview_full.npz    a-4 one step at 64x1024, B=A=3: strided samples + checksums
  crossview_full_trans.npz  a-5 one step at 64x1024, V=A=8 (BASELINE configs 3/4: configured offsets, the shipped existTotal
                        mask, inpainting and rows-0::4 densification masks, settings 7 and 8): strided samples + checksums
"""
import argparse
import os
import sys

import numpy as np

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/LiDARGen")

import torch  # noqa: E402

torch.Tensor.cuda = lambda self, *a, **k: self     # ncsnv2.py:495 calls .cuda() unconditionally

import models as ref_models  # noqa: E402
from models.KITTISampling import anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti as ref_pose  # noqa: E402
from models.ncsnv2 import NCSN_LiDAR_small  # noqa: E402

from oracle.weights import make_state_dict  # noqa: E402
from tests.golden import cases  # noqa: E402

N = argparse.Namespace


class Recorder:
    """Wrap torch entry points the reference calls so that internals can be read back."""

    def __init__(self, noise_list=None):
        self.rounds, self.sparse = [], []
        self.noise = list(noise_list) if noise_list is not None else None
        self._orig = {}

    def __enter__(self):
        self._orig = dict(round=torch.round, sparse=torch.sparse_coo_tensor, randn_like=torch.randn_like)
        rec = self

        def round_(t, *a, **k):
            out = rec._orig["round"](t, *a, **k)
            rec.rounds.append(out.clone())
            return out

        def sparse_(idx, val, *a, **k):
            rec.sparse.append((idx.clone(), val.clone()))
            return rec._orig["sparse"](idx, val, *a, **k)

        def randn_like_(t, *a, **k):
            return rec.noise.pop(0).to(t.dtype)

        torch.round = round_
        torch.sparse_coo_tensor = sparse_
        if self.noise is not None:
            torch.randn_like = randn_like_
        return self

    def __exit__(self, *exc):
        torch.round = self._orig["round"]
        torch.sparse_coo_tensor = self._orig["sparse"]
        torch.randn_like = self._orig["randn_like"]


def dense(idx, val, R, W):
    return torch.sparse_coo_tensor(idx, val, size=(R, W)).to_dense()


def run_one_step(kind, case, sigma, setting):
    """One step with sharing enabled (minStepToShare=0), single level whose sigma == sigmas[-1]."""
    A = case["A"]
    sig = np.array([sigma], dtype=np.float32)
    x = case["x"].clone()
    zero_score = lambda xx, yy: torch.zeros_like(xx)
    with Recorder(noise_list=[torch.zeros_like(x)]) as rec:
        if kind == "pose":
            images, _, shared = ref_pose(x, case["refer"], case["mask"], case["sky"], None, 0, setting,
                                         case["allowance"], zero_score, sig, case["fromWorld"], case["toWorld"],
                                         A, n_steps_each=1, step_lr=0.0, existMask=case["exist"], denoise=False,
                                         verbose=False, grad_ref=0.0, correlation_coefficient=case["coef"])
        else:
            images, _, shared = ref_models.anneal_Langevin_dynamics_inpainting_simultaneous_basic(
                x, case["refer"], case["mask"], case["sky"], None, 0, setting, zero_score, sig,
                case["mods"], A, n_steps_each=1, step_lr=0.0, existMask=case["exist"], denoise=False,
                verbose=False, grad_ref=0.0, correlation_coefficient=case["coef"])
    B, H, W, R = case["B"], case["H"], case["W"], case["R"]
    # the level is both c==0 (sharedImages) and the last level (images[0]); images[-1] is final x
    new_images = images[0]
    x_final = images[-1]
    colr, rowr = rec.rounds[0], rec.rounds[1]                 # pre-flip rounded doubles [B, A*HW]
    per_origin = len(rec.sparse) // B
    cnt = torch.zeros(B, R, W, dtype=torch.int64)
    sum_d = torch.zeros(B, R, W, dtype=torch.float64)
    sum_i = torch.zeros(B, R, W, dtype=torch.float32)
    min_d = torch.zeros(B, R, W, dtype=torch.float64)
    min_i = torch.zeros(B, R, W, dtype=torch.float32)
    have_min = per_origin == 6
    for t in range(B):
        calls = rec.sparse[t * per_origin:(t + 1) * per_origin]
        sum_d[t] = dense(*calls[0], R, W)
        sum_i[t] = dense(*calls[1], R, W)
        cnt[t] = dense(*calls[2], R, W)
        if have_min:
            min_d[t] = dense(*calls[3], R, W)
            min_i[t] = dense(*calls[4], R, W)
    res = dict(new_images=new_images.numpy(), x_final=x_final.numpy(),
               colr=colr.numpy().astype(np.int32), rowr=rowr.numpy().astype(np.int32),
               cnt=cnt.numpy().astype(np.int32), sum_d=sum_d.numpy(), sum_i=sum_i.numpy())
    if have_min:
        res.update(min_d=min_d.numpy(), min_i=min_i.numpy())
    return res


def save(name, **arrs):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrs)
    print(f"wrote {name}: {os.path.getsize(path) / 1024:.1f} KiB")


def gen_sigmas():
    out = {}
    for L in (232, 10):
        cfg = N(model=N(sigma_dist="geometric", sigma_begin=50, sigma_end=0.01, num_classes=L), device="cpu")
        out[f"geometric_{L}"] = ref_models.get_sigmas(cfg).numpy()
    save("sigmas.npz", **out)


def ref_config(H, W, num_classes):
    return N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
             model=N(ngf=128, num_classes=num_classes, nonlinearity="elu", normalization="InstanceNorm++",
                     sigma_dist="geometric", sigma_begin=50, sigma_end=0.01, spec_norm=False),
             device="cpu")


def gen_scorenet():
    H, W, L = 16, 64, 232
    net = NCSN_LiDAR_small(ref_config(H, W, L))
    sd = make_state_dict(ngf=128, num_classes=L, seed=1234)
    net.load_state_dict(sd, strict=True)
    net.eval()
    x, y = cases.scorenet_input(H, W)
    taps = {}
    names = ["begin_conv", "res1.0", "res1.1", "res2.0", "res2.1", "res3.0", "res3.1", "res4.0", "res4.1",
             "refine1", "refine2", "refine3", "refine4"]
    hooks = []
    for nme in names:
        mod = net
        for part in nme.split("."):
            mod = mod[int(part)] if part.isdigit() else getattr(mod, part)
        hooks.append(mod.register_forward_hook(lambda m, i, o, nme=nme: taps.__setitem__(nme, o.detach().clone())))
    with torch.no_grad():
        out = net(x, y)
    for h in hooks:
        h.remove()
    arrs = dict(x=x.numpy(), y=y.numpy(), out=out.numpy())
    for nme, t in taps.items():
        arrs["tap:" + nme] = cases.subsample_tap(t).numpy()
    save("scorenet_small.npz", **arrs)


def gen_crossview():
    case = cases.small_multiview(kind="pose")
    arrs = {}
    for tag, (sigma, setting) in {"hi": (7.5, 5), "lo": (0.3, 5), "lo_nofilter": (0.3, 1)}.items():
        r = run_one_step("pose", case, sigma, setting)
        arrs.update({f"{tag}:{k}": v for k, v in r.items()})
        arrs[f"{tag}:sigma"] = np.float32(sigma)
        arrs[f"{tag}:setting"] = np.int32(setting)
    save("crossview_pose.npz", **arrs)

    case = cases.small_multiview(kind="trans")
    arrs = {}
    for tag, (sigma, setting) in {"hi7": (7.5, 7), "lo7": (0.3, 7), "lo4": (0.3, 4), "lo8": (0.3, 8)}.items():
        r = run_one_step("trans", case, sigma, setting)
        arrs.update({f"{tag}:{k}": v for k, v in r.items()})
        arrs[f"{tag}:sigma"] = np.float32(sigma)
        arrs[f"{tag}:setting"] = np.int32(setting)
    save("crossview_trans.npz", **arrs)

    case = cases.small_multiview(kind="pose", outlier=True)
    r = run_one_step("pose", case, 0.3, 5)
    save("crossview_toohigh.npz", **r)


def gen_samplers():
    sig = cases.short_sigmas()
    # a-4
    case = cases.small_multiview(kind="pose")
    score = cases.fake_score(sig)
    noise = cases.noise_list(case["x"].shape, len(sig) * 2, seed=77)
    with Recorder(noise_list=noise):
        images, _, shared = ref_pose(case["x"].clone(), case["refer"], case["mask"], case["sky"], None, 1, 5,
                                     case["allowance"], score, sig, case["fromWorld"], case["toWorld"], case["A"],
                                     n_steps_each=2, step_lr=6.2e-6, existMask=case["exist"], denoise=True,
                                     verbose=False, grad_ref=1, correlation_coefficient=0.01)
    save("sampler_pose.npz", n_images=np.int32(len(images)), n_shared=np.int32(len(shared)),
         **{f"images{i}": t.numpy() for i, t in enumerate(images)},
         **{f"shared{i}": t.numpy() for i, t in enumerate(shared)})
    # a-5
    case = cases.small_multiview(kind="trans")
    noise = cases.noise_list(case["x"].shape, len(sig) * 2, seed=78)
    with Recorder(noise_list=noise):
        images, _, shared = ref_models.anneal_Langevin_dynamics_inpainting_simultaneous_basic(
            case["x"].clone(), case["refer"], case["mask"], case["sky"], None, 1, 7, score, sig, case["mods"],
            case["A"], n_steps_each=2, step_lr=6.2e-6, existMask=case["exist"], denoise=True, verbose=False,
            grad_ref=1, correlation_coefficient=0.01)
    save("sampler_trans.npz", n_images=np.int32(len(images)), n_shared=np.int32(len(shared)),
         **{f"images{i}": t.numpy() for i, t in enumerate(images)},
         **{f"shared{i}": t.numpy() for i, t in enumerate(shared)})
    # a-6
    noise = cases.noise_list(case["x"].shape, len(sig) * 2, seed=79)
    with Recorder(noise_list=noise):
        images, targets = ref_models.anneal_Langevin_dynamics_inpainting(
            case["x"].clone(), case["refer"], case["mask"], score, sig, n_steps_each=2, step_lr=6.2e-6,
            denoise=True, verbose=False, grad_ref=1)
    save("sampler_single.npz", n_images=np.int32(len(images)),
         **{f"images{i}": t.numpy() for i, t in enumerate(images)})


def gen_n4():
    """row N4: the unconditional and the beam-densification samplers (models/__init__.py:20-109)."""
    sig = cases.short_sigmas()
    case = cases.small_multiview(kind="trans")
    score = cases.fake_score(sig)
    noise = cases.noise_list(case["x"].shape, len(sig) * 2, seed=80)
    with Recorder(noise_list=noise):
        images = ref_models.anneal_Langevin_dynamics(case["x"].clone(), score, sig, n_steps_each=2, step_lr=6.2e-6,
                                                     final_only=False, verbose=False, denoise=True)
    arrs = {"u_n": np.int32(len(images))}
    arrs.update({f"u{i}": t.numpy() for i, t in enumerate(images)})
    noise = cases.noise_list(case["x"].shape, len(sig) * 2, seed=81)
    with Recorder(noise_list=noise):
        images, targets = ref_models.anneal_Langevin_dynamics_densification(
            case["x"].clone(), case["refer"], score, sig, n_steps_each=2, step_lr=6.2e-6, denoise=True, verbose=False,
            grad_ref=0.1, sampling_step=4)
    arrs["d_n"] = np.int32(len(images))
    arrs.update({f"d{i}": t.numpy() for i, t in enumerate(images)})
    save("sampler_n4.npz", **arrs)


def gen_full():
    case = cases.full_multiview()
    r = run_one_step("pose", case, 0.3, 5)
    s = cases.FULL_STRIDE
    flat = lambda a: a.reshape(-1)
    save("crossview_full.npz",
         new_images_s=flat(r["new_images"])[::s], x_final_s=flat(r["x_final"])[::s],
         colr_s=flat(r["colr"])[::s], rowr_s=flat(r["rowr"])[::s], cnt_s=flat(r["cnt"])[::s],
         min_d_s=flat(r["min_d"])[::s], min_i_s=flat(r["min_i"])[::s],
         colr_sum=np.int64(r["colr"].astype(np.int64).sum()), rowr_sum=np.int64(r["rowr"].astype(np.int64).sum()),
         colr_wsum=np.int64((r["colr"].astype(np.int64).reshape(-1) * (np.arange(r["colr"].size) % 1009)).sum()),
         rowr_wsum=np.int64((r["rowr"].astype(np.int64).reshape(-1) * (np.arange(r["rowr"].size) % 1009)).sum()),
         cnt_sum=np.int64(r["cnt"].sum()), n_filled=np.int64((r["cnt"] > 0).sum()),
         new_images_abs_sum=np.float64(np.abs(r["new_images"].astype(np.float64)).sum()),
         min_d_sum=np.float64(r["min_d"].sum()))


def _full_summary(r):
    s = cases.FULL_STRIDE
    flat = lambda a: a.reshape(-1)
    w = np.arange(r["colr"].size) % 1009
    return dict(
        new_images_s=flat(r["new_images"])[::s], x_final_s=flat(r["x_final"])[::s],
        colr_s=flat(r["colr"])[::s], rowr_s=flat(r["rowr"])[::s], cnt_s=flat(r["cnt"])[::s],
        colr_sum=np.int64(r["colr"].astype(np.int64).sum()), rowr_sum=np.int64(r["rowr"].astype(np.int64).sum()),
        colr_wsum=np.int64((r["colr"].astype(np.int64).reshape(-1) * w).sum()),
        rowr_wsum=np.int64((r["rowr"].astype(np.int64).reshape(-1) * w).sum()),
        cnt_sum=np.int64(r["cnt"].sum()), n_filled=np.int64((r["cnt"] > 0).sum()),
        new_images_abs_sum=np.float64(np.abs(r["new_images"].astype(np.float64)).sum()),
        x_final_abs_sum=np.float64(np.abs(r["x_final"].astype(np.float64)).sum()))


def gen_full_trans():
    arrs = {}
    for tag, (sigma, setting, densify) in cases.FULL_TRANS_RUNS.items():
        case = cases.full_translation(densify=densify)
        r = run_one_step("trans", case, sigma, setting)
        out = _full_summary(r)
        if "min_d" in r:
            s = cases.FULL_STRIDE
            out.update(min_d_s=r["min_d"].reshape(-1)[::s], min_i_s=r["min_i"].reshape(-1)[::s],
                       min_d_sum=np.float64(r["min_d"].sum()))
        arrs.update({f"{tag}:{k}": v for k, v in out.items()})
        print(tag, "filled cells", int(out["n_filled"]), "candidates", int(out["cnt_sum"]))
    save("crossview_full_trans.npz", **arrs)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ["sigmas", "scorenet", "crossview", "samplers", "n4", "full", "full_trans"]
    for w in which:
        {"sigmas": gen_sigmas, "scorenet": gen_scorenet, "crossview": gen_crossview,
         "samplers": gen_samplers, "n4": gen_n4, "full": gen_full, "full_trans": gen_full_trans}[w]()
