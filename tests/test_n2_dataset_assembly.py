"""Row N2: multi-view dataset assembly (calibration chain, scan re-rendering from another pose, input
post-processing).  CPU: the numpy oracle against the fixture produced by executing the reference's own source lines.
GPU: the CUDA path (C ABI / host mirror, incl. the file-backed dataset) against the oracle and the fixture."""
import os

import numpy as np
import pytest

from oracle import dataset_assembly_ref as da
from tests.golden import cases

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "dataset_assembly.npz"))
H, W = cases.N2_SHAPE
ITEMS = (1, 5, 8)


def _item_inputs(idx, frames, table):
    view, pose_num = idx % cases.N2_BATCH, idx // cases.N2_BATCH
    wanted = min(pose_num + (view + 1) * 5, len(frames) - 1)
    f0, f1 = frames[pose_num], frames[wanted]
    return cases.n2_scan(f0), cases.n2_scan(f1), table[f0], table[f1], int(f0)


def test_oracle_pose_chain_and_items_match_reference_golden():
    cam_to_velo, cam_to_pose, poses = cases.n2_calibration()
    frames, table = da.pose_chain(cam_to_velo, cam_to_pose[0], poses)          # the reference keeps row 0 (:53)
    assert np.array_equal(frames, G["frames"])
    assert np.array_equal(np.stack([table[f] for f in frames]), G["poses"])
    for idx in ITEMS:
        scan, goal, t0, t1, f0 = _item_inputs(idx, frames, table)
        r = da.assemble_view(scan, goal, t0, t1, H, W)
        t = f"i{idx}:"
        assert f0 == int(G[t + "scan"])
        assert np.array_equal(r["toWorld"], G[t + "toWorld"]) and np.array_equal(r["fromWorld"], G[t + "fromWorld"])
        assert np.array_equal(r["toOGView"], G[t + "toOGView"])
        assert np.array_equal(r["real"], G[t + "real"]) and np.array_equal(r["goalDepth"], G[t + "goal"])
        assert np.array_equal(np.packbits(r["known"]), G[t + "known"])
        assert np.array_equal(np.packbits(r["notsky"]), G[t + "notsky"]) and r["notsky"].all()
        # the synthetic scans repeat 50 points exactly: the reference breaks such depth ties arbitrarily (SURVEY 8a, quirk xii)
        assert int((r["index"].astype(np.int32) != G[t + "index"]).sum()) <= 50
        assert 0.2 < r["known"][0].mean() < 0.98                       # holes and obfuscated pixels exist


def test_host_pose_chain_matches_reference_golden():
    """the calibration chain is host code (float64 numpy): it must reproduce the reference's matrices bit for bit, no GPU needed"""
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import datasets
    cam_to_velo, cam_to_pose, poses = cases.n2_calibration()
    frames, table = datasets.velo_to_world_poses(cam_to_velo, cam_to_pose[0], poses)
    assert np.array_equal(frames, G["frames"]) and np.array_equal(np.stack([table[f] for f in frames]), G["poses"])


@pytest.mark.gpu
def test_cuda_assembly_matches_oracle_and_golden(tmp_path):
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import datasets
    cam_to_velo, cam_to_pose, poses = cases.n2_calibration()
    frames, table = datasets.velo_to_world_poses(cam_to_velo, cam_to_pose[0], poses)
    assert np.array_equal(frames, G["frames"]) and np.array_equal(np.stack([table[f] for f in frames]), G["poses"])
    for idx in ITEMS:
        scan, goal, t0, t1, f0 = _item_inputs(idx, frames, table)
        real, known, notsky, index, toW, fromW, goalDepth, toOG = datasets.assemble_view(scan, goal, t0, t1, rowMax=H, colMax=W)
        ref = da.assemble_view(scan, goal, t0, t1, H, W)
        t = f"i{idx}:"
        assert np.array_equal(toW, G[t + "toWorld"]) and np.array_equal(fromW, G[t + "fromWorld"]) and np.array_equal(toOG, G[t + "toOGView"])
        # CUDA's float64 log2 and numpy's differ by an ulp on ~1 pixel in 8 (values within 4e-16); the 4x4 products and atan2
        # may also differ in the last ulp, so a point on a rounding boundary can move to the neighbouring pixel: count
        # those and bound them (N1 uses the same bound).
        diff = np.abs(real[0] - ref["real"][0]) > 1e-12
        ulp = int((real[0] != ref["real"][0]).sum())
        print(f"[N2 item {idx}] range pixels differing from the numpy oracle: {int(diff.sum())} of {diff.size} ({ulp} by an ulp of log2)")
        assert int(diff.sum()) <= 8
        same = ~diff
        assert np.allclose(real[0][same], G[t + "real"][0][same], rtol=0, atol=4e-16)
        assert np.array_equal(real[1][same], ref["real"][1][same])
        assert int((known != ref["known"]).sum()) <= 16 and notsky.all() and notsky.shape == (1, H, W)
        assert int((np.abs(goalDepth[0] - ref["goalDepth"][0]) > 1e-12).sum()) <= 8
        assert known.shape == (2, H, W) and known.dtype == bool and np.array_equal(known[0], known[1])
        assert index.shape == (1, H, W)
    # file-backed dataset with the reference's directory layout
    root = tmp_path / "KITTI-360"
    drive = "2013_05_28_drive_0000_sync"
    (root / "calibration").mkdir(parents=True)
    (root / "data_poses" / drive).mkdir(parents=True)
    data = root / "data_3d_raw" / drive / "velodyne_points" / "data"
    data.mkdir(parents=True)
    np.savetxt(root / "calibration" / "calib_cam_to_velo.txt", cam_to_velo[None])
    np.savetxt(root / "calibration" / "calib_cam_to_pose.txt", cam_to_pose)
    np.savetxt(root / "data_poses" / drive / "poses.txt", poses)
    for f in frames:
        cases.n2_scan(f).tofile(data / (str(int(f)).zfill(10) + ".bin"))
    import argparse
    NS = argparse.Namespace
    cfg = NS(data=NS(channels=2, image_size=H, image_width=W), sampling=NS(actualBatchSize=cases.N2_BATCH))
    ds = datasets.KITTI360Line(str(root), cfg)
    assert len(ds) == len(frames) * cases.N2_BATCH
    item = ds[5]
    assert item[-1] == int(G["i5:scan"]) and np.array_equal(item[4], G["i5:toWorld"])
    assert int((np.abs(item[0][0] - G["i5:real"][0]) > 1e-12).sum()) <= 8
    with pytest.raises(RuntimeError):
        ds.load_scan(99999)


def test_item_batches_collate_like_the_dataloader():
    """CPU: `ItemBatches.batch(i)` stacks consecutive items into the tuple the runners unpack (stub dataset: the CUDA
    assembly itself is covered by the gpu test above)."""
    import torch
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import datasets

    class Stub:
        def __len__(self):
            return 7

        def __getitem__(self, i):
            f = np.full
            return (f((2, 4, 8), float(i)), f((2, 4, 8), i % 2 == 0), f((1, 4, 8), True), f((1, 4, 8), float(i)),
                    f((1, 4, 4), float(i)), f((1, 4, 4), -float(i)), f((2, 4, 8), 0.5 * i), f((4, 4), float(i)), 100 + i)

    b = datasets.ItemBatches(Stub(), 3)
    assert len(b) == 2
    real, known, notsky, index, to_w, from_w, goal, to_og, frames = b.batch(1)
    assert real.shape == (3, 2, 4, 8) and real.dtype == torch.float64 and known.dtype == torch.bool
    assert notsky.shape == (3, 1, 4, 8) and index.shape == (3, 1, 4, 8) and goal.shape == (3, 2, 4, 8)
    assert to_w.shape == (3, 1, 4, 4) and from_w.shape == (3, 1, 4, 4) and to_og.shape == (3, 4, 4)
    assert frames.tolist() == [103, 104, 105] and real[:, 0, 0, 0].tolist() == [3.0, 4.0, 5.0]
    with pytest.raises(IndexError):
        b.batch(2)


def test_synthetic_batches_have_the_dataset_tuple_layout():
    """CPU: the synthetic stand-in yields the tuple the reference datasets return after the DataLoader's collate
    (kitti360_im_8Batch.py:299-304), i.e. the same layout `ItemBatches` produces from file-backed items"""
    import torch
    import sdpc_b200  # noqa: F401
    from sdpc_b200.synthetic_data import SyntheticMultiView
    B, A, Hs, Ws = 6, 3, 16, 64
    for mode in ("line", "allforone", "densification"):
        real, mask, sky, index, to_w, from_w, goal, to_og, frames = SyntheticMultiView(Hs, Ws, B, A, mode=mode, seed=3).batch(1)
        assert real.shape == (B, 2, Hs, Ws) and real.dtype == torch.float64 and goal.shape == real.shape
        assert mask.shape == (B, 2, Hs, Ws) and mask.dtype == torch.bool and torch.equal(mask[:, 0], mask[:, 1])
        assert sky.shape == (B, 1, Hs, Ws) and sky.dtype == torch.bool and bool(sky.all())
        assert index.shape == (B, 1, Hs, Ws)
        assert to_w.shape == (B, 1, 4, 4) and from_w.shape == (B, 1, 4, 4) and to_w.dtype == torch.float64
        assert to_og.shape == (B, 4, 4) and frames.shape == (B,)
        eye = torch.eye(4, dtype=torch.float64).expand(B, 4, 4)
        assert torch.allclose(torch.matmul(from_w[:, 0], to_w[:, 0]), eye, atol=1e-12)
        assert float(real.min()) >= 0.0 and float(real[:, 0].max()) <= 1.0
        if mode == "densification":                                                              # target keeps every 4th beam
            assert not bool(mask[0, 0, 1].any()) and bool(mask[0, 0, 0].any()) and bool(mask[1, 0, 1].any())
