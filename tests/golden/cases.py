"""Deterministic inputs shared by make_golden.py (reference side) and the tests
(oracle / CUDA side).  numpy PCG64 streams only, so they regenerate bit-identically
on any box.  Shapes follow the dataset tuple of
/root/reference/LiDARGen/datasets/kitti360_im_8Batch.py:304 (real [B,2,H,W], mask,
sky [B,1,H,W], toWorld/fromWorld [B,1,4,4] float64)."""
import numpy as np
import torch

FULL_STRIDE = 487     # sample stride for the 64x1024 fixture


def _rng(*seed):
    return np.random.Generator(np.random.PCG64(list(seed)))


def big_rows(H):
    return int(50 * H // 28)


def line_poses(B, A, step=5.0, yaw=0.01, lateral=0.3):
    """B views in groups of A, each group a line of sensor poses along +x with a small yaw."""
    to_world = np.zeros((B, 1, 4, 4), dtype=np.float64)
    for b in range(B):
        i = b % A
        g = b // A
        a = yaw * i + 0.05 * g
        T = np.eye(4)
        T[:3, :3] = [[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]]
        T[:3, 3] = [step * (i + 1), lateral * i - 0.2 * g, 0.05 * i]
        to_world[b, 0] = T
    from_world = np.linalg.inv(to_world)
    return torch.from_numpy(to_world), torch.from_numpy(from_world)


def smooth_range_image(B, H, W, seed):
    """log-range in [0,1]: d ~ 2..60 m, smoothed along rows; intensity U(0,0.5)."""
    r = _rng(seed, 1)
    d = r.uniform(2.0, 60.0, size=(B, H, W))
    k = np.ones(5) / 5
    d = np.apply_along_axis(lambda v: np.convolve(np.concatenate([v[-2:], v, v[:2]]), k, mode="valid"), 2, d)
    depth = np.clip(np.log2(d + 1) / 6, 0, 1)
    inten = r.uniform(0, 0.5, size=(B, H, W))
    return torch.from_numpy(np.stack([depth, inten], 1).astype(np.float32))


def small_multiview(kind="pose", B=4, A=2, H=16, W=64, seed=2024, outlier=False):
    r = _rng(seed, 2)
    refer = smooth_range_image(B, H, W, seed)
    # sample: reference + noise, with a share of negative ranges to exercise the mirror path
    x = refer + torch.from_numpy(r.normal(0, 0.15, size=refer.shape).astype(np.float32))
    neg = torch.from_numpy(r.uniform(size=(B, H, W)) < 0.08)
    x[:, 0] = torch.where(neg, -x[:, 0].abs() * 0.5, x[:, 0])
    if outlier:
        x[1, 0, 3, 5] = 200.0
    mask = torch.from_numpy((r.uniform(size=(B, 1, H, W)) < 0.6).astype(np.int32)).repeat(1, 2, 1, 1).contiguous()
    sky = torch.ones(B, 1, H, W, dtype=torch.bool)
    if kind == "trans":        # the API allows False sky pixels even if the KITTI datasets never produce them
        sky = torch.from_numpy(r.uniform(size=(B, 1, H, W)) < 0.9)
    exist = torch.from_numpy(r.uniform(size=(A, H, W)) < 0.85)
    case = dict(B=B, A=A, H=H, W=W, R=big_rows(H), x=x, refer=refer, mask=mask, sky=sky, exist=exist,
                allowance=10, coef=0.25)
    if kind == "pose":
        case["toWorld"], case["fromWorld"] = line_poses(B, A, step=1.5, yaw=0.03)
    else:
        case["mods"] = torch.tensor([[0, 0, 0], [5, -5, 0], [-5, -5, 0], [0, 5, 0], [-10, 10, 0], [10, 10, 0],
                                     [-10, 0, 0]], dtype=torch.int64)
    return case


def full_multiview(B=3, A=3, H=64, W=1024, seed=4242):
    r = _rng(seed, 3)
    refer = smooth_range_image(B, H, W, seed)
    x = refer + torch.from_numpy(r.normal(0, 0.05, size=refer.shape).astype(np.float32))
    neg = torch.from_numpy(r.uniform(size=(B, H, W)) < 0.03)
    x[:, 0] = torch.where(neg, -x[:, 0].abs() * 0.5, x[:, 0])
    mask = torch.from_numpy((r.uniform(size=(B, 1, H, W)) < 0.6).astype(np.int32)).repeat(1, 2, 1, 1).contiguous()
    sky = torch.ones(B, 1, H, W, dtype=torch.bool)
    exist = torch.from_numpy(r.uniform(size=(1, H, W)) < 0.7).repeat(A, 1, 1).contiguous()
    to_world, from_world = line_poses(B, A, step=5.0, yaw=0.01)
    return dict(B=B, A=A, H=H, W=W, R=big_rows(H), x=x, refer=refer, mask=mask, sky=sky, exist=exist,
                allowance=10, coef=0.01, toWorld=to_world, fromWorld=from_world)


CONFIG3_MODIFICATIONS = [[0, 0, 0], [5, -5, 0], [-5, -5, 0], [0, 5, 0], [-10, 10, 0], [10, 10, 0], [-10, 0, 0], [10, 0, 0]]
FULL_TRANS_RUNS = {"hi7": (7.5, 7, False), "lo8": (0.3, 8, False), "lo7d": (0.3, 7, True)}   # tag: sigma, setting, densify


def full_translation(densify=False, B=8, A=8, H=64, W=1024, seed=5151):
    """a-5 at the full image size in the shape of BASELINE configs 3 / 4: V = A = 8 views at the configured
    offsets (Inpainting.yml's seven + one more), existMask = the reference's own processed data file (kept bit-packed in
    the package), view 0 the target; `densify`: its known pixels are rows 0::4 only (16 of 64 beams).  A tenth of the sky
    flags are cleared (the KITTI datasets never do that, the API allows it), 3 % of the ranges are negative."""
    import sdpc_b200  # noqa: F401
    from sdpc_b200.synthetic_data import lidargen_exist_mask
    r = _rng(seed, 6)
    refer = smooth_range_image(B, H, W, seed)
    x = refer + torch.from_numpy(r.normal(0, 0.05, size=refer.shape).astype(np.float32))
    neg = torch.from_numpy(r.uniform(size=(B, H, W)) < 0.03)
    x[:, 0] = torch.where(neg, -x[:, 0].abs() * 0.5, x[:, 0])
    known = r.uniform(size=(B, 1, H, W)) < 0.6
    supp = r.uniform(size=(B, 1, H, W)) < 0.9
    if densify:
        rows = np.zeros((1, 1, H, 1), dtype=bool)
        rows[:, :, 0::4] = True
        known = np.broadcast_to(rows, known.shape)
    first = (np.arange(B) % A == 0).reshape(B, 1, 1, 1)
    mask = torch.from_numpy(np.ascontiguousarray(np.where(first, known, supp)).astype(np.int32)).repeat(1, 2, 1, 1).contiguous()
    sky = torch.from_numpy(r.uniform(size=(B, 1, H, W)) < 0.9)
    ex = lidargen_exist_mask(H, W)
    assert ex is not None, "data/exist_mask_lidargen.npz is missing (tests/golden/make_golden_exist.py writes it)"
    exist = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(ex, (A, H, W))))
    return dict(B=B, A=A, H=H, W=W, R=big_rows(H), x=x, refer=refer, mask=mask, sky=sky, exist=exist, allowance=10,
                coef=0.01, mods=torch.tensor(CONFIG3_MODIFICATIONS[:A], dtype=torch.int64))


def short_sigmas():
    """4-level schedule spanning sigma>1 (sigmaMod=sigma) and sigma<=1 (sigmaMod=1); numpy float32
    like the runner's get_sigmas(config).cpu().numpy() (ncsn_runner_kitti_simultaneous.py:491-492)."""
    return np.array([4.0, 1.3, 0.2, 0.01], dtype=np.float32)


def fake_score(sigmas):
    """cheap deterministic stand-in for the score net: pulls towards a smooth field."""
    sig = torch.from_numpy(np.asarray(sigmas, dtype=np.float32))

    def score(x, y):
        H, W = x.shape[-2:]
        target = 0.5 + 0.25 * torch.sin(torch.arange(W, dtype=torch.float32, device=x.device) * (6.283185 / W)).view(1, 1, 1, W)
        return -(x - target) / (sig.to(x.device)[y].view(-1, 1, 1, 1) ** 2) * 0.05
    return score


def noise_list(shape, n, seed):
    r = _rng(seed, 4)
    return [torch.from_numpy(r.standard_normal(size=tuple(shape)).astype(np.float32)) for _ in range(n)]


def scorenet_input(H, W, B=2, seed=99):
    r = _rng(seed, 5)
    x = torch.from_numpy(r.uniform(0, 1, size=(B, 2, H, W)).astype(np.float32))
    y = torch.tensor([3, 200][:B], dtype=torch.int64)
    return x, y


def subsample_tap(t):
    """[B,C,H,W] -> strided sample (every 16th channel, 4th row, 8th column)."""
    return t[:, ::16, ::4, ::8].contiguous()


# ---- row N1: point cloud -> range image -------------------------------------------------------------------
N1_CASES = {"small": (6000, 16, 64, 11), "full": (60000, 64, 1024, 12)}       # tag: (points, H, W, seed)


def synthetic_scan(n, seed):
    """[n,4] float64 points (x, y, z, remission) of a ground plane + walls seen from near the origin, plus clutter
    (several points per pixel, points outside the vertical field of view, a few exactly repeated points)."""
    r = _rng(seed, 9)
    az = r.uniform(-np.pi, np.pi, n)
    el = np.radians(r.uniform(-27.0, 5.0, n))
    d_ground = np.where(np.sin(el) < -0.02, 1.73 / np.maximum(-np.sin(el), 1e-3), 80.0)
    d_wall = 25.0 / np.maximum(np.abs(np.cos(az)) * np.cos(el), 0.05)
    d = np.minimum(np.minimum(d_ground, d_wall), 70.0) * r.uniform(0.97, 1.03, n)
    d = np.where(r.uniform(size=n) < 0.05, d * r.uniform(0.2, 0.9, n), d)      # foreground clutter
    pts = np.stack([d * np.cos(az) * np.cos(el), d * np.sin(az) * np.cos(el), d * np.sin(el), r.uniform(0, 1, n)], 1)
    pts[-50:] = pts[:50]                                                       # exact duplicates (ties)
    origin = np.array([0.3, -0.2, 0.1])
    return pts, origin


# ---- row N2: synthetic KITTI-360-like drive (calibration, poses, scans) ------------------------------------------------
N2_SHAPE = (32, 256)           # rowMax, colMax of the fixture
N2_FRAMES = 24
N2_BATCH = 3                   # actualBatchSize: views per item group
N2_MODIFICATIONS = [[0, 0, 0], [5, -5, 0], [-5, -5, 0]]     # config.data.modifications (first rows of Inpainting.yml)


def n2_calibration():
    """(cam_to_velo 12, cam_to_pose rows [k,12], poses [F,13]) with the layouts of calib_cam_to_velo.txt,
    calib_cam_to_pose.txt (after stripping the labels) and poses.txt."""
    def rigid(yaw, pitch, t):
        cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
        R = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]]) @ np.array([[cp, 0, sp], [0, 1, 0], [-sp, 0, cp]])
        return np.concatenate((R, np.asarray(t, dtype=np.float64).reshape(3, 1)), 1).reshape(-1)

    cam_to_velo = rigid(0.02, -0.01, [0.8, 0.3, -0.6])
    cam_to_pose = np.stack([rigid(0.05, 0.02, [1.6, 0.06, 1.3]), rigid(0.0, 0.0, [0, 0, 0])])
    poses = []
    for f in range(N2_FRAMES):                       # a gentle left curve, ~1 m per frame; frame ids skip some numbers
        yaw = 0.01 * f
        poses.append(np.concatenate(([1 + f + (f // 7)], rigid(yaw, 0.002 * f, [1.0 * f, 0.02 * f * f, 0.01 * f]))))
    return cam_to_velo, cam_to_pose, np.asarray(poses)


def n2_scan(frame, n=20000):
    """float32 [n,4] raw scan of a frame, like np.fromfile(<frame>.bin, float32).reshape(-1, 4)"""
    pts, _ = synthetic_scan(n, 1000 + int(frame))
    return pts.astype(np.float32)
