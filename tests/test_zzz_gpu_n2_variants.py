"""GPU: the AllForOne and densification datasets of row N2 through the C ABI / host mirror, against the numpy
oracle and the fixture made from the reference's own source lines.  Same bounds as tests/test_n2_dataset_assembly.py: CUDA's
float64 log2 / atan2 may differ from numpy's in the last ulp, so a point on a rounding boundary can land in the neighbouring
pixel; such pixels are counted, printed and bounded.  Collected last (written in a session without box minutes: the kernels
behind it are the validated ones of rows N1 / N2, the composition is new)."""
import numpy as np
import pytest

from oracle import dataset_assembly_ref as da
from tests.golden import cases
from tests.test_n2_dataset_assembly import AFO_ITEMS, DEN_ITEMS, GV, H, MODS, W, _write_drive

pytestmark = pytest.mark.gpu


def _compare(item, ref, tag, moved_bound):
    real, known, notsky, index, toW, fromW, goal, toOG = item
    assert np.array_equal(toW, GV[tag + "toWorld"]) and np.array_equal(fromW, GV[tag + "fromWorld"])
    assert np.array_equal(toOG, GV[tag + "toOGView"])
    assert real.shape == (2, H, W) and known.shape == (2, H, W) and known.dtype == bool and np.array_equal(known[0], known[1])
    assert notsky.shape == (1, H, W) and notsky.all() and index.shape == (1, H, W) and goal.shape == (2, H, W)
    diff = np.abs(real[0] - ref["real"][0]) > 1e-12
    print(f"[N2 {tag}] range pixels differing from the numpy oracle: {int(diff.sum())} of {diff.size}; "
          f"known-mask pixels: {int((known != ref['known']).sum())}")
    assert int(diff.sum()) <= moved_bound
    same = ~diff
    assert np.allclose(real[0][same], GV[tag + "real"][0][same], rtol=0, atol=4e-16)
    assert np.array_equal(real[1][same], ref["real"][1][same])
    assert int((known != ref["known"]).sum()) <= 2 * moved_bound
    assert int((np.abs(goal[0] - ref["goalDepth"][0]) > 1e-12).sum()) <= 8


def test_cuda_allforone_and_densification_match_oracle_and_golden(tmp_path):
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import datasets
    cam_to_velo, cam_to_pose, poses = cases.n2_calibration()
    frames, table = datasets.velo_to_world_poses(cam_to_velo, cam_to_pose[0], poses)
    for idx in AFO_ITEMS:
        view, pose = idx % cases.N2_BATCH, idx // cases.N2_BATCH
        f0, f1 = frames[pose], frames[da.allforone_selection(pose, len(frames))]
        scan, goal = cases.n2_scan(f0), cases.n2_scan(f1)
        item = datasets.assemble_view(scan, goal, table[f0], table[f1], MODS[view], rowMax=H, colMax=W)
        _compare(item, da.assemble_view(scan, goal, table[f0], table[f1], H, W, origin=MODS[view]), f"afo{idx}:", 8)
    for idx in DEN_ITEMS:
        view, pose = idx % cases.N2_BATCH, idx // cases.N2_BATCH
        f0 = frames[pose]
        scan = cases.n2_scan(f0)
        item = datasets.assemble_densification_view(scan, table[f0], MODS, view, rowMax=H, colMax=W)
        ref = da.assemble_view_densification(scan, table[f0], MODS, view, H, W)
        _compare(item, ref, f"den{idx}:", 16)                      # two projections in a row: twice the bound
        if view == 0:
            assert not item[1][:, :, :W // 4].any() and item[1][:, :, W // 4:].all()
            assert float(item[0][0][:, :W // 4].max()) <= np.log2(1.0001) / 6 + 1e-12      # blanked quarter holds no point
    # file-backed readers with the reference's directory layout
    cfg = _write_drive(tmp_path / "KITTI-360")
    afo = datasets.KITTI360AllForOne(str(tmp_path / "KITTI-360"), cfg)
    den = datasets.KITTI360Densification(str(tmp_path / "KITTI-360"), cfg)
    a, d = afo[8], den[4]
    assert a[-1] == int(GV["afo8:scan"]) and np.array_equal(a[5], GV["afo8:fromWorld"])
    assert int((np.abs(a[0][0] - GV["afo8:real"][0]) > 1e-12).sum()) <= 8
    assert d[-1] == int(GV["den4:scan"]) and np.array_equal(d[4], GV["den4:toWorld"])
    assert int((np.abs(d[0][0] - GV["den4:real"][0]) > 1e-12).sum()) <= 16
    batch = datasets.ItemBatches(den, cases.N2_BATCH).batch(0)
    assert batch[0].shape == (cases.N2_BATCH, 2, H, W) and batch[7].shape == (cases.N2_BATCH, 4, 4)
