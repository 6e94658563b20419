"""CPU: the C-ABI library loads and exports every symbol include/sdpc_b200.h declares."""
import os
import re

import sdpc_b200  # noqa: F401
from sdpc_b200 import cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    import __graft_entry__ as g
    g.build()


def test_library_exports_every_declared_symbol():
    _build()
    lib = cabi.load()
    header = open(os.path.join(ROOT, "include", "sdpc_b200.h")).read()
    declared = set(re.findall(r"\b(sdpc_[a-z0-9_]+)\s*\(", header))
    bound = {name for name, _, _ in cabi.SYMBOLS}
    assert declared == bound, (declared ^ bound)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.sdpc_abi_version() == cabi.ABI_VERSION == int(re.search(r"#define SDPC_ABI_VERSION (\d+)", header).group(1))
    assert lib.sdpc_build_arch() == b"sm_100a"


def test_struct_layouts_match_header():
    import ctypes
    import subprocess
    import tempfile
    src = '#include "include/sdpc_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(sdpc_step_params), sizeof(sdpc_step_buffers), sizeof(sdpc_score_config), offsetof(sdpc_step_params, allowance), offsetof(sdpc_step_buffers, dbg_min_d));}'
    d = tempfile.mkdtemp()
    open(os.path.join(d, "s.c"), "w").write(src)
    subprocess.check_call(["gcc", "-I", ROOT, os.path.join(d, "s.c"), "-o", os.path.join(d, "s")], cwd=ROOT)
    out = subprocess.check_output([os.path.join(d, "s")]).split()
    assert [int(v) for v in out] == [ctypes.sizeof(cabi.StepParams), ctypes.sizeof(cabi.StepBuffers),
                                     ctypes.sizeof(cabi.ScoreConfig), cabi.StepParams.allowance.offset,
                                     cabi.StepBuffers.dbg_min_d.offset]


def test_product_path_fails_loudly_without_cuda():
    import pytest
    import torch
    from sdpc_b200.samplers import anneal_Langevin_dynamics_inpainting
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    x = torch.zeros(1, 2, 16, 64)
    with pytest.raises(cabi.SdpcError):
        anneal_Langevin_dynamics_inpainting(x, x, torch.zeros_like(x).int(), None, [1.0])


def test_step_kernel_launches():
    """host-only query behind bench.py's gpu_launches: update + (scatter, resolve, re-arm, fix, correct) - the z-buffers are re-armed
    by the resolve pass, so a step has no memset besides the 4-byte max word - in every winner mode."""
    import ctypes as C
    lib = cabi.load()
    p, b = cabi.StepParams(), cabi.StepBuffers()
    p.height, p.width, p.share = 64, 1024, 1
    for mode in (0, 1, 2):
        p.winner_mode = mode
        assert lib.sdpc_step_kernel_launches(C.byref(p), C.byref(b)) == 6
    p.share = 0
    assert lib.sdpc_step_kernel_launches(C.byref(p), C.byref(b)) == 1
    assert lib.sdpc_step_kernel_launches(None, None) < 0


def test_struct_sizes_are_checked_at_load_time():
    """a library built from another revision of the header (different struct layout) must not load silently"""
    import ctypes as C
    lib = cabi.load()
    for which, st in enumerate((cabi.StepParams, cabi.StepBuffers, cabi.ScoreConfig, cabi.ProjectionParams)):
        assert lib.sdpc_abi_struct_bytes(which) == C.sizeof(st)
    assert lib.sdpc_abi_struct_bytes(99) == 0
