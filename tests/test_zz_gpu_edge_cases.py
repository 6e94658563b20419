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


@pytest.mark.parametrize("kind,shape", [("pose", (4, 16, 64)), ("pose", (3, 64, 1024)), ("trans", (4, 16, 64))])
def test_step_stays_inside_its_buffers(kind, shape):
    """compute-sanitizer is closed on this pool, so bounds are checked with canaries: the sample, newImages, the gradient of
    the likelihood and the step workspace live inside larger allocations filled with a pattern; after steps in every
    winner mode (flagged cells and the fix pass included) every byte outside the declared extents is untouched."""
    import ctypes as C
    from sdpc_b200 import cabi
    from tests.test_gpu_crossview import _runner
    B, H, W = shape
    case = cases.small_multiview(kind) if (H, W) == (16, 64) else cases.full_multiview(B=B, A=B)
    dev = "cuda:0"
    guard = 4096                                                   # floats / bytes on either side
    n = case["x"].numel()

    def guarded_f32():
        buf = torch.full((n + 2 * guard,), 12345.678, device=dev)
        return buf, buf[guard:guard + n].view(case["x"].shape)

    xb, x = guarded_f32()
    nb, ni = guarded_f32()
    gb, gl = guarded_f32()
    run = _runner(kind, case, debug=False)
    ws_bytes = run.workspace.numel()
    wsb = torch.full((ws_bytes + 2 * guard,), 0xA5, dtype=torch.uint8, device=dev)
    run.workspace = wsb[guard:guard + ws_bytes]
    assert run.workspace.data_ptr() % 256 == 0
    cabi.check(run.lib, run.lib.sdpc_step_workspace_init(C.c_void_p(run.workspace.data_ptr()), ws_bytes, case["B"], case["H"],
                                                         case["W"], run.geo.R, run._stream()), "init")
    grad = torch.randn(case["x"].shape, device=dev)
    noise = torch.randn(case["x"].shape, device=dev)
    for mode, shift in ((0, 0), (1, 0), (2, 0), (0, 50), (1, 50)):
        x.copy_(case["x"].to(dev))
        run.winner_mode, run.key_shift_override = mode, shift
        if kind == "pose":
            p = run.params(1e-6, 1e-3, 1.0, case["coef"], 1, True, True, 10.0, False)
        else:
            p = run.params(1e-6, 1e-3, 1.0, case["coef"], 1, True, True, 10.0, True)
        run.step(p, run.buffers(x, grad, noise, grad_likelihood=gl, new_images=ni))
    torch.cuda.synchronize()
    for name, buf in (("x", xb), ("new_images", nb), ("grad_likelihood", gb)):
        assert bool((buf[:guard] == 12345.678).all()) and bool((buf[guard + n:] == 12345.678).all()), name
    assert bool((wsb[:guard] == 0xA5).all()) and bool((wsb[guard + ws_bytes:] == 0xA5).all())
    assert bool(torch.isfinite(ni).all()) and float(ni.abs().max()) > 0
