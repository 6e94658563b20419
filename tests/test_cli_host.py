"""CPU: the re-hosted CLI boundary (LiDARGen/main.py:17-163,177-209) - flags, YAML keys, forced overrides, output folder,
and the reference's error convention (an exception while sampling is logged with its traceback, the process returns 0)."""
import logging
import os

import pytest
import torch
import yaml

import sdpc_b200  # noqa: F401
from sdpc_b200 import main as cli

CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(sdpc_b200.__file__)), "configs")


@pytest.mark.parametrize("name,dataset", [("Line.yml", "KITTI360_im_8batch"), ("Inpainting.yml", None), ("Densification.yml", None)])
def test_shipped_configs_parse_with_reference_flags(tmp_path, name, dataset):
    exp = str(tmp_path / "exp")
    args, cfg = cli.parse_args_and_config(["--sample", "--ni", "--config", name, "--exp", exp, "--doc", "d", "-i", "imgs",
                                           "--seed", "7"])
    assert args.image_folder == os.path.join(exp, "image_samples", "imgs") and os.path.isdir(args.image_folder)
    assert args.log_path == os.path.join(exp, "logs", "d")
    # forced overrides of the reference (main.py:46-48)
    assert cfg.sampling.inpainting is True and cfg.sampling.interpolation is False and cfg.sampling.densification is False
    # keys the samplers and runners read
    for key in ("batch_size", "actualBatchSize", "n_steps_each", "step_lr", "denoise", "ckpt_id"):
        assert hasattr(cfg.sampling, key), key
    for key in ("channels", "image_size", "image_width", "dataset", "logit_transform", "rescaled"):
        assert hasattr(cfg.data, key), key
    for key in ("ngf", "num_classes", "sigma_begin", "sigma_end", "sigma_dist", "normalization", "nonlinearity", "ema", "ema_rate"):
        assert hasattr(cfg.model, key), key
    assert cfg.sampling.batch_size % cfg.sampling.actualBatchSize == 0
    if dataset:
        assert cfg.data.dataset == dataset
    else:
        assert hasattr(cfg.data, "modifications") and len(cfg.data.modifications) >= cfg.sampling.actualBatchSize
    assert torch.initial_seed() == 7


def test_only_sampling_is_supported(tmp_path):
    with pytest.raises(SystemExit):
        cli.parse_args_and_config(["--ni", "--config", "Line.yml", "--exp", str(tmp_path)])


def test_sampling_errors_are_logged_and_return_zero(tmp_path, caplog):
    """without a CUDA device the samplers raise (no CPU fallback); like the reference's main(), the CLI logs the
    traceback and still returns 0"""
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    cfg = yaml.safe_load(open(os.path.join(CFG_DIR, "Line.yml")))
    cfg["sampling"].update(batch_size=4, actualBatchSize=2, n_steps_each=1)
    cfg["data"].update(image_size=16, image_width=64)
    cfg["model"].update(num_classes=3)
    p = tmp_path / "Line.yml"
    p.write_text(yaml.safe_dump(cfg))
    with caplog.at_level(logging.ERROR):
        assert cli.main(["--sample", "--ni", "--config", str(p), "--exp", str(tmp_path / "exp")]) == 0
    assert "no CPU fallback" in caplog.text or "CUDA" in caplog.text


def test_cli_accepts_every_reference_flag(tmp_path):
    """flags of the reference's main.py (tests/golden/cli_flags.json, read from its source): every one parses here with
    the same action; defaults are the reference's except the three site-specific ones (--config, --exp, --doc)"""
    import json
    flags = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "cli_flags.json")))
    assert {"--sample", "--ni", "--config", "--seed", "--exp", "--doc", "--image_folder"} <= {f["names"][-1] for f in flags}
    argv = ["--config", "Line.yml", "--exp", str(tmp_path), "--ni"]
    for f in flags:
        name = f["names"][-1]
        if f["action"] == "store_true" and name not in ("--ni",):
            argv.append(name)
    args, _ = cli.parse_args_and_config(argv)                           # --sample is among them, so parsing succeeds
    for f in flags:
        dest = f["names"][-1].lstrip("-")
        assert hasattr(args, dest), dest
        if f["action"] == "store_true":
            assert getattr(args, dest) is True, dest
    args, _ = cli.parse_args_and_config(["--config", "Line.yml", "--exp", str(tmp_path), "--ni", "--sample"])
    for f in flags:
        dest = f["names"][-1].lstrip("-")
        if dest in ("config", "exp", "doc", "sample", "ni", "image_folder"):
            continue
        want = False if f["action"] == "store_true" else f["default"]
        assert getattr(args, dest) == want, (dest, getattr(args, dest), want)


@pytest.mark.parametrize("name", ["Line.yml", "Inpainting.yml", "Densification.yml"])
def test_shipped_configs_carry_the_reference_values(name):
    """data / model / sampling sections equal the reference's configuration files key for key (HDVMine_Line.yml,
    HDVMine_Circle.yml, HDVMine_Densification.yml; tests/golden/reference_configs.json); only the `b200` section is new"""
    import json
    ref = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_configs.json")))[name]
    ours = yaml.safe_load(open(os.path.join(CFG_DIR, name)))
    for sec in ("data", "model", "sampling"):
        assert ours[sec] == ref[sec], (name, sec)
    assert set(ours) - {"data", "model", "sampling"} == {"b200"}


@pytest.mark.parametrize("name,arms", [("Line.yml", 3), ("Inpainting.yml", 3), ("Densification.yml", 2)])
def test_runner_flow_with_stand_in_samplers(tmp_path, monkeypatch, name, arms):
    """CPU dry run of `runner.sample()`: the three sampler entry points and the model loader are replaced by stand-ins that
    check what they are called with and return image lists of the right shapes, so the ablation loop (`doThis`), the view
    selection per group and every output file of a batch are exercised without a GPU
    (ncsn_runner_kitti_simultaneous.py:527-893, ncsn_runner_AllForOne.py:540-994)"""
    import glob
    import numpy as np
    from sdpc_b200 import runner
    cfg = yaml.safe_load(open(os.path.join(CFG_DIR, name)))
    cfg["sampling"].update(batch_size=6, actualBatchSize=3, n_steps_each=1)
    cfg["data"].update(image_size=16, image_width=64)
    cfg["model"].update(num_classes=4)
    cfg["b200"].update(max_batches=1)
    p = tmp_path / name
    p.write_text(yaml.safe_dump(cfg))
    calls = []

    def images(x):
        return [torch.full(tuple(x.shape), 0.25).reshape(-1), torch.full(tuple(x.shape), 0.75).reshape(-1)]

    def pose(x, refer, mask, sky, idx, start, setting, allowance, score, sigmas, fromW, toW, A, n_steps, lr, existMask=None,
             denoise=True, verbose=True, grad_ref=0.1, correlation_coefficient=0.1, sampling_step=16):
        assert x.shape == refer.shape == mask.shape and sky.shape[0] == x.shape[0] and x.shape[0] % A == 0
        assert tuple(fromW.shape) == (x.shape[0], 4, 4) and tuple(toW.shape) == (x.shape[0], 4, 4)
        assert (start, setting, allowance, grad_ref, correlation_coefficient) == (2, 5, 10, 1, 0.01) and len(sigmas) == 4
        calls.append(("pose", x.shape[0], A))
        return images(x), [], []

    def trans(x, refer, mask, sky, idx, start, setting, score, sigmas, mods, A, n_steps, lr, existMask=None, denoise=True,
              verbose=True, grad_ref=0.1, correlation_coefficient=0.1, sampling_step=16):
        assert x.shape == refer.shape == mask.shape and x.shape[0] % A == 0 and mods.shape[0] >= A and mods.shape[1] == 3
        assert (start, setting, grad_ref, correlation_coefficient) == (2, 7, 1, 0.01)
        calls.append(("trans", x.shape[0], A))
        return images(x), [], []

    def single(x, refer, mask, score, sigmas, n_steps, lr, denoise=True, verbose=True, grad_ref=0.1, sampling_step=16):
        assert x.shape == refer.shape == mask.shape and grad_ref == 1
        calls.append(("single", x.shape[0], 1))
        return images(x), []

    monkeypatch.setattr(runner, "anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti", pose)
    monkeypatch.setattr(runner, "anneal_Langevin_dynamics_inpainting_simultaneous_basic", trans)
    monkeypatch.setattr(runner, "anneal_Langevin_dynamics_inpainting", single)
    monkeypatch.setattr(runner._Base, "load_score", lambda self: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)      # main.py then configures cuda:0; nothing below touches it
    real_rand = torch.rand
    monkeypatch.setattr(torch, "rand", lambda *a, **k: real_rand(*a, **{**k, "device": "cpu"}))
    exp = str(tmp_path / "exp")
    args, config = cli.parse_args_and_config(["--sample", "--ni", "--config", str(p), "--exp", exp, "-i", "out"])
    config.device = torch.device("cpu")
    cls = runner.NCSNRunnerKITTISimultaneous if name == "Line.yml" else runner.NCSNRunnerAllForOne
    assert cls(args, config).sample() == 0
    if name == "Line.yml":            # 2 views of each group, then all 3, then the single-view baseline on all 6
        assert calls == [("pose", 4, 2), ("pose", 6, 3), ("single", 6, 1)]
    elif name == "Inpainting.yml":    # same arms, but the baseline only runs on view 0 of each group
        assert calls == [("trans", 4, 2), ("trans", 6, 3), ("single", 2, 1)]
    else:                             # densification: the full group, then the baseline
        assert calls == [("trans", 6, 3), ("single", 2, 1)]
    out = os.path.join(exp, "image_samples", "out")
    count = lambda pat: len(glob.glob(os.path.join(out, pat)))
    assert count("*_Masked_completion_897.pth.npy") == arms and count("*_TimeTaken.npy") == arms
    assert count("*_Masked_image_grid_897.png") == arms
    assert count("0_*_Input_completion_897.pth.npy") == 1 and count("0_*_GT_completion_897.pth.npy") == 1
    assert count("0_*_SKY_897.pth.npy") == 1 and count("*_Input_image_grid_897.png") == 1 and count("*_GT_image_grid_897.png") == 1
    assert count("toWorld_*.npy") == 1 and count("fromWorld_*.npy") == 1
    assert count("*_Shared_completion_initial897.pth.npy") == (0 if name == "Line.yml" else arms)
    for f in glob.glob(os.path.join(out, "*_Masked_completion_897.pth.npy")):
        a = np.load(f)
        n = [c for c in calls][int(os.path.basename(f)[0])][1]
        assert a.shape == (2 * n, 3, 16, 64) and float(a.min()) == 0.75 and float(a.max()) == 0.75
    for f in glob.glob(os.path.join(out, "*_Shared_completion_initial897.pth.npy")):
        assert float(np.load(f).max()) == 0.25


def test_bench_reference_arm_prints_the_contract_line():
    """CPU: `bench.py --impl reference` (the arm the driver times beside the GPU arm) runs the oracle port on one view-step of
    the B = A = 8 workload and prints ONE JSON line with the contract's keys; a non-zero rank prints nothing."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, check=True).stdout.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    assert line["impl"] == "reference" and line["metric"] == "view-steps/sec" and line["steps"] == 1 and line["warmup"] == 0
    assert line["value"] > 0 and line["higher_is_better"] is True and line["cpu_baseline"]["kind"] == "port"
    assert line["config"]["group_size"] == 8 and line["e2e"]["h2d_bytes_per_step"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    quiet = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                           capture_output=True, text=True, timeout=600, check=True, env=env).stdout.strip()
    assert quiet == ""


def test_bench_workloads_and_traffic_record():
    """CPU: the benchmark's input generator (BASELINE configs 2-5) and the ncu traffic record bench.py's roofline reads."""
    import json
    import sys
    import numpy as np
    import torch
    from sdpc_b200 import build as b
    from sdpc_b200.synthetic_data import INPAINTING_MODIFICATIONS, bench_group, lidargen_exist_mask
    ex = lidargen_exist_mask(64, 1024)
    assert ex is not None and ex.shape == (64, 1024) and abs(ex.mean() - 0.680) < 5e-4 and int(ex.any(axis=1).sum()) == 57
    assert lidargen_exist_mask(16, 64) is None
    for variant in ("line", "inpainting", "densification"):
        g = bench_group(16, 8, 64, 1024, 1234, variant)
        g2 = bench_group(16, 8, 64, 1024, 1234, variant)
        assert all(torch.equal(g[k], g2[k]) for k in g if isinstance(g[k], torch.Tensor))          # reproducible bytes
        assert g["x"].shape == (16, 2, 64, 1024) and g["mask"].dtype == torch.int32 and g["exist"].shape == (8, 64, 1024)
        assert torch.equal(g["exist"][0], torch.from_numpy(ex)) and bool(g["sky"].all())
        if variant == "line":
            assert g["toWorld"].shape == (16, 1, 4, 4) and g["toWorld"].dtype == torch.float64
            eye = torch.bmm(g["toWorld"].reshape(16, 4, 4), g["fromWorld"].reshape(16, 4, 4))
            assert torch.allclose(eye, torch.eye(4, dtype=torch.float64).expand(16, 4, 4), atol=1e-9)
        else:
            assert g["mods"].tolist() == INPAINTING_MODIFICATIONS[:8]
            first, others = g["mask"][0::8, 0].float().mean().item(), g["mask"][1, 0].float().mean().item()
            assert abs(others - 0.9) < 0.01
            if variant == "densification":                       # the target keeps rows 0::4 only: 16 of 64 beams
                assert abs(first - 0.25) < 1e-6 and bool((g["mask"][0, 0, 0::4] == 1).all()) and int(g["mask"][0, 0, 1::4].sum()) == 0
            else:
                assert abs(first - 0.6) < 0.01
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    rec = json.load(open(os.path.join(root, "profiles", "conv_traffic.json")))
    assert set(rec["arms"]) >= {"bf16", "bf16x3", "fp16"} and all(a["n"] >= 4 for a in rec["arms"].values())
    sys.path.insert(0, root)
    import bench
    traffic, src = bench.measured_conv_traffic("bf16")
    if rec["kernel_source_digest"] == b.kernel_digest():        # the committed capture is of the kernels in this tree
        assert traffic == rec["arms"]["bf16"]["dram_bytes_per_launch_mean"] and 5e7 < traffic < 1e9
    else:                                                        # sources moved on: the figure is dropped, never guessed
        assert traffic is None and "older build" in src
    assert bench.measured_conv_traffic("fp32")[0] is None
