"""Golden vectors for the evaluation error sums of row N3, from the reference's own notebook cell.

Cell 1 of MeasureResults/QuantifyingNotebookSynthesis_Line.ipynb walks the .npy files of a finished Line run (42 views in
6 groups of 7, settings 0..6) and accumulates the L1 depth / intensity errors per (setting, view-in-group).  It cannot
run on the reference's data (absent) nor import cv2 here, so this script
  * writes a synthetic run directory with the reference runner's file names and array layouts (seeded float32 images,
    regenerated bit-identically by `case_arrays` in the tests),
  * executes THE CELL'S OWN SOURCE (read from /root/reference at generation time) with a stand-in `cv2` whose
    inpaint / resize only feed the classical-baseline totals this repo does not restate,
  * stores the totals the cell leaves behind.
Run in the build container only; the fixture travels, /root/reference does not.

    python tests/golden/make_golden_n3_errors.py
"""
import contextlib
import glob
import io
import json
import os
import shutil
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
NOTEBOOK = "/root/reference/MeasureResults/QuantifyingNotebookSynthesis_Line.ipynb"
BATCH, GROUP, H, W = 42, 7, 64, 1024          # batchNum / actualBatchNum of the cell, image size of its reshapes
SETTINGS = 7                                    # largestSettingNum = 6 in the cell


def views_kept(setting):
    """numberOfOutputs of the cell: views per group a setting's result file holds."""
    return min(setting + 2, GROUP)


def case_arrays(seed=21):
    """gt, inp [42,2,64,1024] float32 and one prediction per setting [6*k,2,64,1024] float32."""
    rng = np.random.default_rng(seed)
    gt = rng.uniform(0.0, 1.0, size=(BATCH, 2, H, W)).astype(np.float32)
    gt[:, 0] = np.where(rng.uniform(size=(BATCH, H, W)) < 0.02, np.float32(1.02), gt[:, 0])   # a few ranges beyond 63 m
    inp = (gt * (rng.uniform(size=gt.shape) < 0.6)).astype(np.float32)
    preds = []
    for s in range(SETTINGS):
        k = views_kept(s)
        idx = np.array([GROUP * g + j for g in range(BATCH // GROUP) for j in range(k)])
        preds.append(np.clip(gt[idx] + rng.normal(0, 0.03, size=(len(idx), 2, H, W)), 0, 1).astype(np.float32))
    return gt, inp, preds


def grid_layout(x):
    """[B,2,H,W] -> [2B,3,H,W], the layout the runners save (ncsn_runner_kitti_simultaneous.py:848-870)."""
    x = np.transpose(x, (1, 0, 2, 3)).reshape(-1, 1, H, W)
    return np.concatenate((x, x, x), 1)


def write_run(folder, gt, inp, preds):
    save_num = "".join(str(7 * g) + "_" for g in range(BATCH // GROUP))          # the runner's saveNum
    tag = "_897.pth.npy"
    open(os.path.join(folder, "0_0_GT_image_grid_897.png"), "wb").close()        # only its name is read
    np.save(os.path.join(folder, "0_" + save_num + "_TimeTaken.npy"), np.float64(1.0))
    np.save(os.path.join(folder, "0_" + save_num + "_GT_completion" + tag[:-4]), grid_layout(gt))
    np.save(os.path.join(folder, "0_" + save_num + "_Input_completion" + tag[:-4]), grid_layout(inp))
    for s, p in enumerate(preds):
        np.save(os.path.join(folder, str(s) + "_" + save_num + "_Masked_completion" + tag[:-4]), grid_layout(p))
    rng = np.random.default_rng(5)
    poses = np.tile(np.eye(4), (BATCH, 1, 1, 1))
    poses[:, 0, :3, 3] = rng.uniform(-20, 20, size=(BATCH, 3))
    np.save(os.path.join(folder, "toWorld_" + save_num), poses)
    np.save(os.path.join(folder, "fromWorld_" + save_num), np.linalg.inv(poses))


def fake_cv2():
    m = types.ModuleType("cv2")
    m.INTER_NEAREST, m.INTER_LINEAR, m.INTER_CUBIC = 0, 1, 2
    m.inpaint = lambda img, mask, radius, flags=0: np.array(img, copy=True)
    m.resize = lambda src, dsize, *a, fx=1.0, fy=1.0, interpolation=0: np.repeat(np.asarray(src), int(fy), axis=0)
    return m


def run_cell(folder):
    nb = json.load(open(NOTEBOOK))
    src = "".join(nb["cells"][1]["source"])
    env = {"np": np, "os": os, "glob": glob, "cv2": fake_cv2(), "folderOne": folder, "folderTwo": ""}
    with contextlib.redirect_stdout(io.StringIO()):
        exec(src, env)
    return env


def main():
    gt, inp, preds = case_arrays()
    folder = tempfile.mkdtemp(prefix="n3_errors_")
    try:
        write_run(folder, gt, inp, preds)
        env = run_cell(folder)
    finally:
        shutil.rmtree(folder, ignore_errors=True)
    keys = ("totalDistanceError", "totalIntensityError", "totalDistanceErrorInput", "totalIntensityErrorInput",
            "totalDistanceInput", "totalPixels", "totalInputPixels")
    np.savez_compressed(os.path.join(HERE, "n3_errors.npz"), seed=21, **{k: np.asarray(env[k]) for k in keys})
    print({k: np.asarray(env[k]).shape for k in keys})
    print("totalDistanceError[0]:", env["totalDistanceError"][0])


if __name__ == "__main__":
    main()
