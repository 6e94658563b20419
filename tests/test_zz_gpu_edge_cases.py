"""GPU: degenerate inputs of the cross-view step through the C ABI (the same cases run on the host emulation in
tests/test_host_emul.py).  Collected last so that the parity suites above it always run first."""
import pytest
import torch

from tests.golden import cases
from tests.test_gpu_crossview import _cuda_step

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("winner_mode", [1, 2])
def test_no_source_pixel_exists(winner_mode):
    """existMask all False: no candidate reaches any z-buffer - the shared images are zero, nothing is corrected"""
    case = cases.small_multiview("pose")
    case["exist"] = torch.zeros_like(case["exist"])
    x, ni, run = _cuda_step("pose", case, 0.3, 5, debug="cells", winner_mode=winner_mode)
    assert int(run.debug["cnt"].abs().sum()) == 0 and int((run.debug["winner"] != -1).sum()) == 0
    assert float(ni.abs().max()) == 0.0
    assert torch.equal(x.cpu(), case["x"]) and int(run.too_high.item()) == 0


@pytest.mark.parametrize("winner_mode", [1, 2])
def test_every_pixel_known(winner_mode):
    """refer_mask all ones: the shared images do not depend on the mask, the correction (1 - mask) vanishes"""
    case = cases.small_multiview("pose")
    _, ni_ref, _ = _cuda_step("pose", case, 0.3, 5, debug="cells", winner_mode=winner_mode)
    case["mask"] = torch.ones_like(case["mask"])
    x, ni, _ = _cuda_step("pose", case, 0.3, 5, debug="cells", winner_mode=winner_mode)
    assert torch.equal(ni, ni_ref) and float(ni.abs().max()) > 0.0
    assert torch.equal(x.cpu(), case["x"])


@pytest.mark.parametrize("shape", [(2, 16, 64), (4, 64, 1024)])
def test_identity_poses_shift_the_group_mean_down_one_row(shape):
    """size-independent property at the small and at the full image size (the reference's hidden invariant):
    with identical poses every view receives the mean of the group's views shifted down by one row; row 0 stays empty"""
    from tests.test_host_emul import _identity_case
    B, H, W = shape
    case = _identity_case(B, H, W)
    x, ni, run = _cuda_step("pose", case, 0.3, 5, debug="cells")
    x, ni = x.cpu(), ni.cpu()
    mean = case["x"].double().mean(0, keepdim=True).float()
    assert float(ni[:, :, 0].abs().max()) == 0.0
    assert torch.allclose(ni[:, :, 1:], mean[:, :, :-1].expand_as(ni[:, :, 1:]), rtol=0, atol=2e-6)
    crop = run.debug["cnt"].cpu()[:, case["R"] - H:]
    assert int(crop[:, 0].sum()) == 0 and bool((crop[:, 1:] == B).all())
    want = case["x"] + case["coef"] * (-(case["x"] - ni))
    want[:, :, 0] = case["x"][:, :, 0]
    assert torch.allclose(x, want, rtol=0, atol=1e-6)
