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
