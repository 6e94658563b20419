"""CPU: the drop-in `models` package.  With `dropin/` on sys.path the reference
runners' own import statements must resolve, and every entry point must keep the reference's parameter names, order and
defaults (the runners pass most arguments positionally); extra parameters may only be appended with defaults.
The reference signatures come from tests/golden/signatures.json (recorded from the unmodified reference package)."""
import json
import os
import subprocess
import sys

import sdpc_b200  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(os.path.dirname(os.path.abspath(sdpc_b200.__file__)), "dropin")

PROBE = r'''
import inspect, json, sys
# the reference runner's own import statements (runners/ncsn_runner_kitti_simultaneous.py:15-27)
from models import (anneal_Langevin_dynamics_inpainting,
                    anneal_Langevin_dynamics_inpainting_simultaneous_basic, get_sigmas)
from models.KITTISampling import anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti
from models.ncsnv2 import NCSN_LiDAR_small
from models.ema import EMAHelper
import importlib
out = {}
for key in json.load(open(sys.argv[1])):
    mod, dotted = key.split(":")
    obj = importlib.import_module(mod)
    for part in dotted.split("."):
        obj = getattr(obj, part)
    out[key] = [[p.name, None if p.default is inspect.Parameter.empty else repr(p.default)]
                for p in inspect.signature(obj).parameters.values()]
out["__file__"] = importlib.import_module("models").__file__
print(json.dumps(out))
'''


def test_reference_imports_resolve_and_signatures_match():
    fixture = os.path.join(ROOT, "tests", "golden", "signatures.json")
    env = dict(os.environ, PYTHONPATH=DROPIN + os.pathsep + os.environ.get("PYTHONPATH", ""))
    res = subprocess.run([sys.executable, "-c", PROBE, fixture], capture_output=True, text=True, env=env, cwd="/", timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    ours = json.loads(res.stdout.strip().splitlines()[-1])
    assert os.path.dirname(ours.pop("__file__")) == os.path.join(DROPIN, "models")        # not some other `models`
    ref = json.load(open(fixture))
    for key, want in ref.items():
        got = ours[key]
        assert got[:len(want)] == want, (key, got, want)            # same names, order and defaults
        for name, default in got[len(want):]:                       # additions (shard, precision, ...) are optional
            assert default is not None, (key, name)


def test_state_dict_layout_matches_reference_module():
    """same keys, same order, same shapes as the unmodified reference module (tests/golden/state_dict_inventory.json), so a
    reference checkpoint loads with strict=True behind DataParallel's 'module.' prefix and EMAHelper walks the same names"""
    import argparse
    import torch
    from sdpc_b200.ema import EMAHelper
    from sdpc_b200.scorenet import NCSN_LiDAR_small
    inv = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_inventory.json")))
    N = argparse.Namespace
    cfg = N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=64, image_width=1024),
            model=N(ngf=128, num_classes=232, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                    sigma_begin=50, sigma_end=0.01, spec_norm=False), device=torch.device("cpu"))
    net = NCSN_LiDAR_small(cfg, precision="bf16")
    assert [[k, list(v.shape)] for k, v in net.state_dict().items()] == inv["state_dict"]
    assert [k for k, _ in net.named_parameters()] == inv["named_parameters"]
    assert len(inv["state_dict"]) == 154
    # a reference-style checkpoint: states[0] = DataParallel state dict, states[-1] = EMA shadow
    g = torch.Generator().manual_seed(0)
    states0 = {"module." + k: torch.randn(shape, generator=g) * 0.01 for k, shape in inv["state_dict"]}
    dp = torch.nn.DataParallel(net)
    assert dp.load_state_dict(states0, strict=True) is not None
    assert torch.equal(net.state_dict()["refine4.output_convs.3_2_conv.weight"], states0["module.refine4.output_convs.3_2_conv.weight"])
    shadow = {k: torch.full(tuple(shape), 0.5) for k, shape in inv["state_dict"] if k in set(inv["named_parameters"])}
    ema = EMAHelper(mu=0.999)
    ema.register(dp)
    assert sorted(ema.state_dict()) == sorted(inv["named_parameters"])
    ema.load_state_dict(shadow)
    ema.ema(dp)
    assert float(net.state_dict()["begin_conv.bias"].mean()) == 0.5
