"""GPU, 2 ranks, NCCL: the view-sharded sampler (dist.ViewShard: per-step MAX all-reduce + in-place all-gather of
the updated planes) reproduces the single-GPU run bit for bit.  Skipped with fewer than 2 GPUs."""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(rank, world, A, shard):
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import samplers
    from tests.golden import cases
    dev = torch.device("cuda", rank)
    case = cases.small_multiview("pose")
    case["exist"] = case["exist"].repeat(2, 1, 1)
    sig = cases.short_sigmas()
    score = cases.fake_score(sig)
    noise = iter([n.to(dev) for n in cases.noise_list(case["x"].shape, 8, 77)])
    orig = torch.randn_like
    torch.randn_like = lambda t, *a, **k: next(noise)
    try:
        to = lambda t: t.to(dev)
        im, _, _ = samplers.anneal_Langevin_dynamics_inpainting_simultaneous_basic_kitti(
            to(case["x"]), to(case["refer"]), to(case["mask"]), to(case["sky"]), None, 1, 5, 10, score, sig,
            case["fromWorld"], case["toWorld"], A, n_steps_each=2, step_lr=6.2e-6, existMask=to(case["exist"]),
            denoise=True, verbose=False, grad_ref=1, correlation_coefficient=0.01, shard=shard)
    finally:
        torch.randn_like = orig
    return [t.numpy() for t in im]


def _worker(rank, world, port, out_dir, one_gpu=False):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    if one_gpu:                                        # both ranks on cuda:0, gloo carries the exchange (NCCL needs 2 devices)
        torch.cuda.set_device(0)
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        torch.cuda.set_device(rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import sdpc_b200  # noqa: F401
    from sdpc_b200.dist import ViewShard
    res = {}
    for A in (2, 4):                                   # A=2: groups stay on one rank; A=4: the group spans both ranks
        res[A] = _run(0 if one_gpu else rank, world, A, ViewShard(4, A))
    if dist.get_rank() == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), **{f"a{A}_{i}": a for A, v in res.items() for i, a in enumerate(v)})
    dist.destroy_process_group()


def test_sharded_nccl_matches_single_gpu():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out_dir = tempfile.mkdtemp(prefix="sdpc_nccl_")
    mp.spawn(_worker, args=(2, port, out_dir), nprocs=2, join=True)
    got = np.load(os.path.join(out_dir, "sharded.npz"))
    for A in (2, 4):
        ref = _run(0, 1, A, None)
        for i, t in enumerate(ref):
            assert np.array_equal(got[f"a{A}_{i}"], t), (A, i)


def test_sharded_two_ranks_on_one_gpu():
    """the same sharded sampler with both ranks on cuda:0 (gloo carries the exchange): the pack / unpack kernels, the target
    ranges of update / share and the max fold run on a one-GPU box too, and must reproduce the single-process run bit for bit"""
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out_dir = tempfile.mkdtemp(prefix="sdpc_gloo_gpu_")
    mp.spawn(_worker, args=(2, port, out_dir, True), nprocs=2, join=True)
    got = np.load(os.path.join(out_dir, "sharded.npz"))
    for A in (2, 4):
        ref = _run(0, 1, A, None)
        for i, t in enumerate(ref):
            assert np.array_equal(got[f"a{A}_{i}"], t), (A, i)
