"""CPU, world_size 2, gloo: the view-sharded sampler (dist.ViewShard) equals the single-process run.

The CUDA kernels are replaced by the g++ host emulation (tests/host_emul), which exports the product
library's entry-point names; the host logic under test (target ranges, MAX all-reduce of the tooHigh
gate, all-gather of the updated planes, sharded scoring, final gather) is the product's own code."""
import ctypes as C
import os
import socket
import subprocess
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _emul_lib(path):
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import cabi
    lib = C.CDLL(path)
    P, I, SZ = C.c_void_p, C.c_int, C.c_size_t
    for name, res, args in cabi.SYMBOLS:
        if name.startswith("sdpc_step") or name.startswith("sdpc_langevin") or name.startswith("sdpc_crossview") \
                or name == "sdpc_last_error":
            if hasattr(lib, name):
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
    return lib


def _run(kind, group_size, shard, lib, outlier=False):
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import samplers
    from tests.golden import cases
    case = cases.small_multiview(kind, outlier=outlier)
    if group_size > case["exist"].shape[0]:
        case["exist"] = case["exist"].repeat(group_size // case["exist"].shape[0], 1, 1)
    sig = cases.short_sigmas()
    score = cases.fake_score(sig)
    noise = iter(cases.noise_list(case["x"].shape, 8, 77))
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: next(noise)
    try:
        if kind == "pose":
            im, _, _ = samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
                case["x"], case["refer"], case["mask"], case["sky"], None, 1, 5, 10, score, sig, case["fromWorld"],
                case["toWorld"], group_size, n_steps_each=2, step_lr=6.2e-6, existMask=case["exist"], denoise=True,
                verbose=False, grad_ref=1, correlation_coefficient=0.01, shard=shard, _lib=lib)
        else:
            im, _, _ = samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic(
                case["x"], case["refer"], case["mask"], case["sky"], None, 1, 7, score, sig, case["mods"], group_size,
                n_steps_each=2, step_lr=6.2e-6, existMask=case["exist"], denoise=True, verbose=False, grad_ref=1,
                correlation_coefficient=0.01, shard=shard, _lib=lib)
    finally:
        torch.randn_like = orig
    return im


def _worker(rank, world, port, libpath, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sdpc_b200  # noqa: F401
    from sdpc_b200.dist import ViewShard
    lib = _emul_lib(libpath)
    res = {}
    # B = 4 views.  A = 2: every group lives on one rank (no gather); A = 4: the group spans both ranks (gather).
    for tag, kind, A in (("pose_a2", "pose", 2), ("trans_a2", "trans", 2), ("pose_a4", "pose", 4)):
        shard = ViewShard(4, A)
        assert shard.needs_gather == (A == 4)
        im = _run(kind, A, shard, lib)
        res[tag] = [t.numpy() for t in im]
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), **{f"{k}_{i}": a for k, v in res.items() for i, a in enumerate(v)})
    dist.destroy_process_group()


@pytest.fixture(scope="module")
def emul_path():
    out = os.path.join(tempfile.mkdtemp(prefix="sdpc_emul_"), "libemul.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(ROOT, "tests", "host_emul", "crossview_host.cpp")])
    return out


def test_sharded_sampler_matches_single_process(emul_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out_dir = tempfile.mkdtemp(prefix="sdpc_gloo_")
    mp.spawn(_worker, args=(2, port, emul_path, out_dir), nprocs=2, join=True)
    got = np.load(os.path.join(out_dir, "sharded.npz"))
    lib = _emul_lib(emul_path)
    for tag, kind, A in (("pose_a2", "pose", 2), ("trans_a2", "trans", 2), ("pose_a4", "pose", 4)):
        ref = _run(kind, A, None, lib)
        for i, t in enumerate(ref):
            assert np.array_equal(got[f"{tag}_{i}"], t.numpy()), (tag, i)


def test_single_process_emulation_matches_reference_golden(emul_path):
    """the same host path, unsharded, against the reference's golden trajectory (CPU division semantics)."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "sampler_pose.npz"))
    im = _run("pose", 2, None, _emul_lib(emul_path))
    assert len(im) == int(g["n_images"])
    for i, t in enumerate(im):
        assert int((np.abs(t.numpy() - g[f"images{i}"]) > 2e-4).sum()) <= 16


def test_samplers_own_their_buffers(emul_path):
    """ownership contract of the reference: the caller's x_mod / refer_image / refer_mask / sky / existMask
    are never written, and every returned image is a fresh CPU float32 tensor that aliases neither the inputs nor the
    other outputs"""
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import samplers
    from tests.golden import cases
    lib = _emul_lib(emul_path)
    case = cases.small_multiview("pose")
    keep = {k: case[k].clone() for k in ("x", "refer", "mask", "sky", "exist", "toWorld", "fromWorld")}
    sig = cases.short_sigmas()
    im, targets, shared = samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
        case["x"], case["refer"], case["mask"], case["sky"], None, 0, 5, 10, cases.fake_score(sig), sig, case["fromWorld"],
        case["toWorld"], 2, n_steps_each=2, step_lr=6.2e-6, existMask=case["exist"], denoise=True, verbose=False,
        grad_ref=1, correlation_coefficient=0.01, _lib=lib)
    for k, v in keep.items():
        assert torch.equal(case[k], v), k
    assert targets == [] and len(shared) == 2 and len(im) == 3        # shared: level 0 x 2 steps; images: last level x 2 + final
    outs = im + shared
    ptrs = {t.data_ptr() for t in outs}
    assert len(ptrs) == len(outs) and case["x"].data_ptr() not in ptrs
    for t in outs:
        assert t.device.type == "cpu" and t.dtype == torch.float32 and tuple(t.shape) == tuple(case["x"].shape)


def test_reference_print_quirk_is_kept(emul_path, capsys):
    """the pose sampler prints its diagnostics at levels 1 and 2 even with verbose=False
    (KITTISampling.py:492-500); the translation sampler does not; both print the `grad_ref:` line after the denoise step"""
    lib = _emul_lib(emul_path)
    _run("pose", 2, None, lib)
    out = capsys.readouterr().out
    assert "level: 1," in out and "level: 2," in out and "level: 0," not in out and "level: 3," not in out
    assert out.count("grad_ref: 1") == 3                              # levels 1, 2 and the line after the denoise step
    _run("trans", 2, None, lib)
    out = capsys.readouterr().out
    assert "level:" not in out and out.count("grad_ref: 1") == 1


def test_single_view_call_is_accepted(emul_path):
    """deliberate deviation (SURVEY.md 8a quirk viii): the reference's pose sampler cannot be called with one view
    (`torch.squeeze` turns its [1,1,4,4] poses into [4,4] and the batched product fails, KITTISampling.py:22-23,185);
    here B = A = 1 runs: the view is re-projected into itself"""
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import samplers
    from tests.golden import cases
    case = cases.small_multiview("pose")
    sig = cases.short_sigmas()
    one = {k: v[:1].clone() for k, v in case.items() if k in ("x", "refer", "mask", "sky", "toWorld", "fromWorld")}
    im, targets, shared = samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
        one["x"], one["refer"], one["mask"], one["sky"], None, 0, 5, 10, cases.fake_score(sig), sig, one["fromWorld"],
        one["toWorld"], 1, n_steps_each=2, step_lr=6.2e-6, existMask=case["exist"], denoise=True, verbose=False,
        grad_ref=1, correlation_coefficient=0.01, _lib=_emul_lib(emul_path))
    assert len(im) == 3 and targets == [] and all(tuple(t.shape) == (1, 2, case["H"], case["W"]) for t in im + shared)
    assert all(bool(torch.isfinite(t).all()) for t in im)
    assert float(shared[0].abs().max()) > 0.0                         # the cross-view block ran and filled pixels
