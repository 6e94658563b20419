"""Signatures of the reference's hot-path entry points, for the drop-in boundary test.

Imports the UNMODIFIED reference `models` package (build container only) and records, for every function / class the
reference runners import from it, the parameter names in order with their defaults - the runners pass most arguments
positionally, so the order is the contract.  The fixture travels, /root/reference does not.

    python tests/golden/make_golden_signatures.py
"""
import inspect
import json
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, "/root/reference/LiDARGen")

WANTED = {
    "models": ["anneal_Langevin_dynamics", "anneal_Langevin_dynamics_densification", "anneal_Langevin_dynamics_inpainting",
               "anneal_Langevin_dynamics_inpainting_simultaneous_basic", "get_sigmas"],
    "models.KITTISampling": ["anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti"],
    "models.ncsnv2": ["NCSN_LiDAR_small.__init__", "NCSN_LiDAR_small.forward"],
    "models.ema": ["EMAHelper.__init__", "EMAHelper.register", "EMAHelper.update", "EMAHelper.ema", "EMAHelper.ema_copy",
                   "EMAHelper.state_dict", "EMAHelper.load_state_dict"],
}


def describe(fn):
    out = []
    for p in inspect.signature(fn).parameters.values():
        out.append([p.name, None if p.default is inspect.Parameter.empty else repr(p.default)])
    return out


def resolve(module, dotted):
    obj = module
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


def main():
    import importlib
    import torch
    if not torch.cuda.is_available():                       # models/ncsnv2.py calls .cuda() at import-free run time only
        pass
    table = {}
    for mod, names in WANTED.items():
        m = importlib.import_module(mod)
        for n in names:
            table[mod + ":" + n] = describe(resolve(m, n))
    with open(os.path.join(HERE, "signatures.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    # state_dict inventory of the reference module: ordered keys with shapes (what a checkpoint's states[0] holds,
    # behind the DataParallel 'module.' prefix)
    import argparse
    N = argparse.Namespace
    cfg = N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=64, image_width=1024),
            model=N(ngf=128, num_classes=232, nonlinearity="elu", normalization="InstanceNorm++", sigma_dist="geometric",
                    sigma_begin=50, sigma_end=0.01, spec_norm=False), device=torch.device("cpu"))
    net = importlib.import_module("models.ncsnv2").NCSN_LiDAR_small(cfg)
    inv = [[k, list(v.shape)] for k, v in net.state_dict().items()]
    named = [k for k, _ in net.named_parameters()]
    with open(os.path.join(HERE, "state_dict_inventory.json"), "w") as f:
        json.dump({"state_dict": inv, "named_parameters": named}, f, indent=0)
    print("state_dict keys", len(inv), "parameters", len(named))
    # command-line flags of LiDARGen/main.py (read from its source: importing it pulls in runners -> h5py)
    import ast
    flags = []
    tree = ast.parse(open("/root/reference/LiDARGen/main.py").read())
    for node in ast.walk(tree):
        if isinstance(node, ast.Call) and getattr(node.func, "attr", "") == "add_argument":
            names = [a.value for a in node.args if isinstance(a, ast.Constant)]
            kw = {k.arg: ast.literal_eval(k.value) for k in node.keywords if k.arg in ("default", "action")}
            flags.append({"names": names, "action": kw.get("action"), "default": kw.get("default")})
    with open(os.path.join(HERE, "cli_flags.json"), "w") as f:
        json.dump(flags, f, indent=1)
    print("cli flags", [f["names"][-1] for f in flags])
    # the data / model / sampling sections of the three configurations the reference README names
    import yaml
    names = {"Line.yml": "HDVMine_Line.yml", "Inpainting.yml": "HDVMine_Circle.yml", "Densification.yml": "HDVMine_Densification.yml"}
    cfgs = {}
    for ours, theirs in names.items():
        c = yaml.safe_load(open(os.path.join("/root/reference/LiDARGen/configs", theirs)))
        cfgs[ours] = {"reference_file": theirs, **{sec: c[sec] for sec in ("data", "model", "sampling")}}
    with open(os.path.join(HERE, "reference_configs.json"), "w") as f:
        json.dump(cfgs, f, indent=1, sort_keys=True)
    print("configs", {k: v["reference_file"] for k, v in cfgs.items()})
    for k, v in table.items():
        print(k, [a for a, _ in v])


if __name__ == "__main__":
    main()
