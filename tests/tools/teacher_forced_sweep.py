"""Teacher-forced parity sweep of the Line.yml sampling step over the WHOLE noise schedule (test infrastructure).

The oracle (torch fp32 on the same GPU, TF32 off) walks the a-4 sampler's trajectory with the RAW random-init score
network: V = A = 8 views at 64x1024, 232 levels, `--steps-each` Langevin steps per level (default 1), minStepToShare 2,
setting 5.  At every step the oracle's own pre-step sample is fed into ONE step of the CUDA path (score forward in the
arm under test, update, cross-view block) with the same injected noise, and four things are recorded:

  score      max-abs error of the CUDA score / max-abs of the oracle's (the arm's tolerance: 1e-3 bf16x3, 8e-2 bf16)
  update     the same for the sample after the Langevin update
  exact      the cross-view block on IDENTICAL input (the oracle's post-update sample): pixel indices, validity, counts,
             winners, nearest depths must be bit-exact against the oracle (asserted), newImages and the corrected
             sample within 1e-6 of the image's largest value
  end to end the CUDA step on its OWN post-update sample against the oracle's step: cells whose count differs (candidates
             that crossed a pixel boundary because the score differs in the last digits - the flip count), and the error
             of the final sample and of newImages outside those cells

The next step starts from the oracle's sample again (teacher forcing), so errors do not accumulate and every level is
tested on the distribution the reference visits.  Writes one JSON document; exits non-zero if an assertion fails.

    python tests/tools/teacher_forced_sweep.py --precision bf16x3 --out gpurun_out/teacher_forced_bf16x3.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import sdpc_b200  # noqa: F401,E402
from sdpc_b200 import cabi  # noqa: E402
from sdpc_b200.scorenet import NCSN_LiDAR_small  # noqa: E402
from sdpc_b200.step import StepRunner  # noqa: E402
from oracle import crossview_ref as cv  # noqa: E402
from oracle import samplers_ref as sr  # noqa: E402
from oracle.scorenet_ref import score_forward  # noqa: E402
from oracle.sigmas import sigma_schedule  # noqa: E402
from oracle.weights import make_state_dict  # noqa: E402
from tests.golden import cases  # noqa: E402

N = argparse.Namespace
TOL = {"fp32": 1e-4, "bf16x3": 1e-3, "fp16": 2e-2, "tf32": 2e-2, "bf16": 8e-2}


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def sweep(precision, V=8, H=64, W=1024, L=232, steps_each=1, levels=None, dev="cuda:0", verbose=True):
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
            model=N(ngf=128, num_classes=L, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                    sigma_begin=50, sigma_end=0.01, spec_norm=False), device=dev)
    sd = make_state_dict(num_classes=L)
    sd_dev = {k: v.to(dev) for k, v in sd.items()}
    sig = sigma_schedule(50, 0.01, L).numpy()
    net = NCSN_LiDAR_small(cfg, precision=precision).to(dev)
    net.load_state_dict(sd)
    case = cases.full_multiview(B=V, A=V)
    to = lambda t: t.to(dev)
    refer, mask, sky, exist = to(case["refer"]), to(case["mask"]), to(case["sky"]), to(case["exist"])
    to_w, from_w = to(case["toWorld"]).squeeze(1), to(case["fromWorld"]).squeeze(1)
    geo = cv.make_geometry(H, W, dev)
    run = StepRunner((V, 2, H, W), dev, refer, mask, sky, exist, V, cabi.SDPC_VARIANT_POSE, to_world=to_w, from_world=from_w,
                     debug="cells")
    run_plain = StepRunner((V, 2, H, W), dev, refer, mask, sky, exist, V, cabi.SDPC_VARIANT_POSE, to_world=to_w,
                           from_world=from_w)
    gen = torch.Generator(device=dev).manual_seed(2024)
    x = torch.rand(V, 2, H, W, device=dev, generator=gen)                  # the runner's x0 (uniform noise)
    min_share, setting, allowance, coef, grad_ref, step_lr = 2, 5, 10.0, 0.01, 1.0, 6.2e-6
    rows, t0 = [], time.time()
    for c in range(L):
        sigma = sig[c]
        sm = sigma if sigma > 1 else 1
        labels = torch.full((V,), c, device=dev, dtype=torch.long)
        step_size, noise_scale = sr._step_constants(step_lr, sigma, sig[-1])
        share = c >= min_share
        for s in range(steps_each):
            measure = levels is None or c in levels
            noise = torch.randn(x.shape, device=dev, generator=gen)
            # ---- oracle step on its own sample
            g_ref = torch.nan_to_num(score_forward(sd_dev, x, labels))
            x_upd, _ = sr.langevin_update(x, g_ref, refer, mask, noise, step_size, noise_scale, grad_ref)
            if share:
                ni_ref, im_ref, th_ref, d = cv.shared_images(x_upd, geo, sm, V, exist, sky, to_world=to_w, from_world=from_w,
                                                             min_depth_filter=True, controlled_average=True,
                                                             allowance=allowance, return_debug=True)
                x_next = cv.apply_correction(x_upd, ni_ref, im_ref, sky, mask, th_ref, coef)
            else:
                x_next = x_upd
            if measure:
                row = dict(level=c, step=s, sigma=float(sigma), share=bool(share))
                # ---- CUDA: score and update on the oracle's pre-step sample
                g_cuda = net(x, labels)
                row["score_rel"] = rel(torch.nan_to_num(g_cuda), g_ref)
                p = run.params(step_size, noise_scale, grad_ref, coef, sm, share, True, allowance, False)
                p_upd = run.params(step_size, noise_scale, grad_ref, coef, sm, False, True, allowance, False)
                x_cu = x.clone()
                run_plain.update_only(p_upd, run_plain.buffers(x_cu, g_cuda, noise))
                row["update_rel"] = rel(x_cu, x_upd)
                if share:
                    # ---- cross-view block on IDENTICAL input: bit-exact integers
                    x_id = x_upd.clone()
                    ni_id = torch.zeros_like(x_id)
                    p0 = run.params(0.0, 0.0, 0.0, coef, sm, True, True, allowance, False)
                    run.too_high.zero_()
                    run.step(p0, run.buffers(x_id, None, None, new_images=ni_id))     # eps = 0: the update is the identity
                    dbg = run.debug
                    ok_cnt = torch.equal(dbg["cnt"], d["cnt"].int())
                    tied = d["n_tied"] > 1
                    ok_win = torch.equal(dbg["winner"][~tied], d["winner"].int()[~tied])
                    ok_min = torch.equal(dbg["min_d"], d["min_d"])
                    row.update(exact_counts=ok_cnt, exact_winners=ok_win, exact_min_depth=ok_min, tied_cells=int(tied.sum()),
                               filled_cells=int((d["cnt"] > 0).sum()), candidates=int(d["cnt"].sum()),
                               too_high=bool(th_ref), new_images_rel_identical_input=rel(ni_id, ni_ref),
                               x_rel_identical_input=rel(x_id, x_next))
                    assert ok_cnt and ok_win and ok_min, row
                    assert int(run.too_high.item()) == int(bool(th_ref)), row
                    # values: against the LARGEST value of the image (at the first levels samples and intensities reach
                    # the hundreds and the oracle's own float32 sums carry 1e-5 absolute rounding on near-cancelling cells)
                    assert row["new_images_rel_identical_input"] <= 1e-6 and row["x_rel_identical_input"] <= 1e-6, row
                    # ---- the whole CUDA step on its own numbers
                    x_e2e = x.clone()
                    ni_e2e = torch.zeros_like(x_e2e)
                    run.step(p, run.buffers(x_e2e, g_cuda, noise, new_images=ni_e2e))
                    flipped = dbg["cnt"] != d["cnt"].int()
                    row["flipped_cells"] = int(flipped.sum())
                    row["flipped_frac"] = row["flipped_cells"] / max(1, row["filled_cells"])
                    row["x_rel"] = rel(x_e2e, x_next)
                    row["new_images_max_abs_diff"] = float((ni_e2e - ni_ref).abs().max())
                    row["new_images_pixels_off_1e-2"] = int(((ni_e2e - ni_ref).abs() > 1e-2 * ni_ref.abs().max()).sum())
                else:
                    row["x_rel"] = row["update_rel"]
                assert row["score_rel"] <= TOL[precision], row
                rows.append(row)
                if verbose and (c % 20 == 0 or c == L - 1):
                    print(f"level {c:3d} sigma {sigma:8.4f}: score {row['score_rel']:.2e} x {row['x_rel']:.2e} "
                          f"flipped {row.get('flipped_cells', 0)} of {row.get('filled_cells', 0)} "
                          f"({time.time() - t0:.0f} s)", flush=True)
            x = x_next                                                       # teacher forcing
    shared = [r for r in rows if r["share"]]
    summary = dict(
        precision=precision, views=V, levels=L, steps_each=steps_each, steps_measured=len(rows),
        score_rel_max=max(r["score_rel"] for r in rows), score_rel_median=float(np.median([r["score_rel"] for r in rows])),
        update_rel_max=max(r["update_rel"] for r in rows), x_rel_max=max(r["x_rel"] for r in rows),
        exact_integer_steps=sum(1 for r in shared if r["exact_counts"] and r["exact_winners"] and r["exact_min_depth"]),
        shared_steps=len(shared), flipped_cells_max=max([r["flipped_cells"] for r in shared] or [0]),
        flipped_frac_max=max([r["flipped_frac"] for r in shared] or [0.0]),
        flipped_cells_total=sum(r["flipped_cells"] for r in shared), filled_cells_total=sum(r["filled_cells"] for r in shared),
        new_images_rel_identical_input_max=max([r["new_images_rel_identical_input"] for r in shared] or [0.0]),
        too_high_steps=sum(1 for r in shared if r["too_high"]), seconds=time.time() - t0, tolerance=TOL[precision])
    return dict(summary=summary, per_step=rows)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--precision", default="bf16x3")
    ap.add_argument("--views", type=int, default=8)
    ap.add_argument("--steps-each", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    res = sweep(a.precision, V=a.views, steps_each=a.steps_each)
    print(json.dumps(res["summary"]))
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            json.dump(res, f, indent=0)
