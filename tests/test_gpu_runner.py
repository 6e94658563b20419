"""GPU: the re-hosted CLI / runners (main.py --sample --config Line.yml|Inpainting.yml|Densification.yml) run end to
end on a reduced configuration and write the reference's output files (names, shapes, value range)."""
import glob
import os

import numpy as np
import pytest
import yaml

import sdpc_b200  # noqa: F401
from sdpc_b200 import main as cli

pytestmark = pytest.mark.gpu
CFG_DIR = os.path.join(os.path.dirname(os.path.abspath(sdpc_b200.__file__)), "configs")


def _reduced(name, tmp_path, **data_over):
    cfg = yaml.safe_load(open(os.path.join(CFG_DIR, name)))
    cfg["sampling"].update(batch_size=6, actualBatchSize=3, n_steps_each=1)
    cfg["data"].update(image_size=16, image_width=64, **data_over)
    cfg["model"].update(num_classes=4)
    cfg["b200"].update(precision="tf32", max_batches=1)
    p = tmp_path / name
    p.write_text(yaml.safe_dump(cfg))
    return str(p)


@pytest.mark.parametrize("name,n_variants", [("Line.yml", 3), ("Inpainting.yml", 3), ("Densification.yml", 2)])
def test_cli_sample_writes_reference_outputs(tmp_path, name, n_variants):
    cfg = _reduced(name, tmp_path)
    exp = str(tmp_path / "exp")
    assert cli.main(["--sample", "--ni", "--config", cfg, "--exp", exp, "-i", "out"]) == 0
    out = os.path.join(exp, "image_samples", "out")
    files = sorted(os.path.basename(f) for f in glob.glob(os.path.join(out, "*")))
    assert any(f.startswith("toWorld_") for f in files) and any(f.startswith("fromWorld_") for f in files)
    masked = sorted(glob.glob(os.path.join(out, "*_Masked_completion_897.pth.npy")))
    times = sorted(glob.glob(os.path.join(out, "*_TimeTaken.npy")))
    assert len(masked) == n_variants and len(times) == n_variants, files
    for f in masked:
        a = np.load(f)
        assert a.ndim == 4 and a.shape[1:] == (3, 16, 64) and a.shape[0] % 2 == 0
        assert np.isfinite(a).all() and a.min() >= 0.0 and a.max() <= 1.0          # inverse_data_transform clamp
    for kind in ("Input_completion_897", "GT_completion_897", "SKY_897"):          # once per batch (doThis = 0)
        assert len(glob.glob(os.path.join(out, "0_*_" + kind + ".pth.npy"))) == 1, (kind, files)
    shared = glob.glob(os.path.join(out, "*_Shared_completion_initial897.pth.npy"))
    assert len(shared) == (0 if name == "Line.yml" else n_variants), files     # only the AllForOne runner writes them
    if name == "Line.yml":
        # doThis = 0: 2 of 3 views per group, 2 groups -> 4 views -> [8,3,16,64]; last variant = baseline on all 6 views
        assert np.load(masked[0]).shape[0] == 8 and np.load(masked[-1]).shape[0] == 12
