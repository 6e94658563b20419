#!/usr/bin/env python
"""bench.py - view-steps/sec of the simultaneous multi-view Langevin sampling step on B200.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.

A step = one Langevin step of the Line.yml sampler (a-4, KITTISampling.py:137-490) over one batch of
synthetic views: score-network forward + noise draw + Langevin update + cross-view block.
Workload (BASELINE.json configs[1]): 8 synthetic line poses per GPU (B = A = 8, one group per rank),
2x64x1024 range/intensity images, random-init NCSN_LiDAR_small (29.7 M parameters), noise level
c = 116 of 232 (sigmaMod = 1, sharing on, setting 5).  Weak scaling: every rank owns one whole group
of 8 views, so the only exchange is the 1-float MAX all-reduce of the tooHigh gate
(KITTISampling.py:162 takes the max over ALL views of the call).

  value    : view-steps/s, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e      : same metric with the step's x copied host->device (pinned) before the score forward and
             x + newImages copied back device->host inside the timed region
  roofline : tensor-core convolutions (the dominant kernel family): algorithmic 2*M*N*K FLOPs of the
             launches in the timed region / their summed CUDA-event time on the launching stream
  cpu_baseline / --impl reference : the oracle port (oracle/*.py, torch CPU ops like the reference)
             on the box's host cores, bounded sample of 1 view (A = 1) per step
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, LEVELS, LEVEL = 64, 1024, 232, 116
VIEWS_PER_GPU = 8
STEP_LR = 6.2e-6


def workload(B=VIEWS_PER_GPU, A=VIEWS_PER_GPU):
    return ("Line.yml step (a-4): NCSN_LiDAR_small forward + Langevin update + cross-view (setting 5, minStepToShare "
            f"passed), {B} line poses per GPU (B={B}, A={A}), 2x64x1024, random-init 29.7M-param net, noise level 116/232")


WORKLOAD = workload()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default=os.environ.get("SDPC_PRECISION", "bf16"), choices=["bf16", "bf16x3", "tf32", "fp32"])
    ap.add_argument("--views-per-gpu", type=int, default=VIEWS_PER_GPU)
    ap.add_argument("--group-size", type=int, default=8,
                    help="actualBatchSize A: views that share information; --views-per-gpu 16/32/64 with the default 8 is "
                         "BASELINE.json's view-count sweep (groups of 8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity-arm", action="store_true", help="skip the extra bf16x3 (fp32-parity) timing")
    return ap.parse_args()


def config_ns(device):
    N = argparse.Namespace
    return N(data=N(logit_transform=False, rescaled=False, channels=2, image_size=H, image_width=W),
             model=N(ngf=128, num_classes=LEVELS, nonlinearity="elu", normalization="InstanceNorm++",
                     sigma_dist="geometric", sigma_begin=50, sigma_end=0.01, spec_norm=False), device=device)


def synthetic_group(B, seed, A=None):
    """B views in groups of A (default: one group): the dataset tuple of kitti360_im_8Batch.py:304 with synthetic content
    (SURVEY 8d)."""
    A = B if A is None else A
    import numpy as np
    import torch
    from tests.golden import cases
    refer = cases.smooth_range_image(B, H, W, seed)
    r = np.random.Generator(np.random.PCG64([seed, 7]))
    mask = torch.from_numpy((r.uniform(size=(B, 1, H, W)) < 0.6).astype(np.int32)).repeat(1, 2, 1, 1).contiguous()
    sky = torch.ones(B, 1, H, W, dtype=torch.bool)
    exist = torch.from_numpy(r.uniform(size=(1, H, W)) < 0.68).repeat(B, 1, 1).contiguous()
    to_world, from_world = cases.line_poses(B, A, step=5.0, yaw=0.01)
    x0 = torch.from_numpy(r.uniform(size=(B, 2, H, W)).astype(np.float32))
    return dict(x=x0, refer=refer, mask=mask, sky=sky, exist=exist, toWorld=to_world, fromWorld=from_world)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1, t_loaded=None):
        """Samples of the timed region [t0, t1]; if the region was shorter than one sampling period, the samples up to
        `t_loaded` (the roofline / end-to-end passes that follow without a pause: same kernels, same load)."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows, window = [r for t, r in self.rows if t0 <= t <= t1], "timed region"
        if not rows and t_loaded is not None:
            rows, window = [r for t, r in self.rows if t0 <= t <= t_loaded], "timed region + the passes that follow it under the same load"
        sm, mx, reasons = [], None, set()
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "window": window}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ---------------------------------------------------------------------------------------------- CPU arm
def cpu_view_steps_per_s(steps, warmup, threads=None):
    """Oracle port on the host cores: one Langevin step (score forward + update + cross-view) of ONE view
    (A = 1) per step - a bounded sample of the 8-view workload; the cross-view cost per view grows with A,
    so this flatters the CPU arm."""
    import numpy as np
    import torch
    from oracle import crossview_ref as cv
    from oracle import samplers_ref as sr
    from oracle.scorenet_ref import score_forward
    from oracle.sigmas import sigma_schedule
    from oracle.weights import make_state_dict
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    sd = make_state_dict(num_classes=LEVELS)
    sig = sigma_schedule(50, 0.01, LEVELS).numpy()
    g = synthetic_group(1, 1234)
    geo = cv.make_geometry(H, W)
    x = g["x"].clone()
    labels = torch.full((1,), LEVEL, dtype=torch.long)
    step_size, noise_scale = sr._step_constants(STEP_LR, sig[LEVEL], sig[-1])
    eye_to, eye_from = g["toWorld"].reshape(1, 4, 4), g["fromWorld"].reshape(1, 4, 4)

    def one(x):
        grad = torch.nan_to_num(score_forward(sd, x, labels))
        noise = torch.randn_like(x)
        x, _ = sr.langevin_update(x, grad, g["refer"], g["mask"], noise, step_size, noise_scale, 1)
        ni, im, th = cv.shared_images(x, geo, 1, 1, g["exist"], g["sky"], to_world=eye_to, from_world=eye_from,
                                      min_depth_filter=True, controlled_average=True, allowance=10.0)
        return cv.apply_correction(x, ni, im, g["sky"], g["mask"], th, 0.01)

    for _ in range(warmup):
        x = one(x)
    t0 = time.perf_counter()
    for _ in range(steps):
        x = one(x)
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps * 1e3, threads


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 8))
    warm = max(1, min(args.warmup, 2))
    v, ms, threads = cpu_view_steps_per_s(steps, warm)
    line = {"impl": "reference", "metric": "view-steps/sec", "value": v, "unit": "view-steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "views_per_gpu": VIEWS_PER_GPU,
                       "cpu_sample": "each step = 1 view (A=1) of the workload: the reference's CPU path needs ~1 s per "
                                     "view-forward, a full 8-view step with its O(A^2) cross-view block ~10 s"},
            "cpu_baseline": {"value": v, "unit": "view-steps/s", "cores": threads, "kind": "port",
                             "sample": f"{steps} steps x 1 view (A=1), oracle port (torch CPU ops), {threads} threads"},
            "e2e": {"value": v, "unit": "view-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------- GPU arm
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import sdpc_b200  # noqa: F401
    from sdpc_b200 import cabi
    from sdpc_b200.scorenet import NCSN_LiDAR_small
    from sdpc_b200.step import StepRunner
    from sdpc_b200.sigmas import get_sigmas

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B = args.views_per_gpu
    A = min(B, args.group_size)
    if B % A:
        raise SystemExit(f"--views-per-gpu {B} is not a multiple of --group-size {A}")
    g = synthetic_group(B, 1234 + rank, A)
    cfg = config_ns(dev)
    sig = get_sigmas(cfg).cpu().numpy()
    torch.manual_seed(1234)                                           # random-init weights (nn.Conv2d-style init of the module)
    net = NCSN_LiDAR_small(cfg, precision=args.precision).to(dev)
    run = StepRunner((B, 2, H, W), dev, g["refer"], g["mask"], g["sky"], g["exist"], A, cabi.SDPC_VARIANT_POSE,
                     to_world=g["toWorld"], from_world=g["fromWorld"])
    x = g["x"].to(dev)
    labels = torch.full((B,), LEVEL, device=dev, dtype=torch.long)
    step_size = STEP_LR * (sig[LEVEL] / sig[-1]) ** 2
    noise_scale = np.sqrt(step_size * 2)
    p = run.params(step_size, noise_scale, 1, 0.01, 1, True, True, 10, False)
    new_images = torch.empty_like(x)
    gmax = torch.zeros(1, device=dev)

    def step(xbuf):
        grad = net(xbuf, labels)
        noise = torch.randn_like(xbuf)
        b = run.buffers(xbuf, grad, noise, new_images=new_images)
        if world == 1:
            run.step(p, b)
        else:                                   # tooHigh is a max over every view of the call
            run.update_only(p, b)
            mx = run.local_max()
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            run.merge_max(mx)
            run.share_only(p, b)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            fn()
        bb.record()
        barrier()
        ms = a.elapsed_time(bb)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- device-resident arm -------------------------------------------------------------------
    clocks = ClockSampler(local) if rank == 0 else None          # started before the warm-up: nvidia-smi needs a moment to stream
    for _ in range(max(args.warmup, 3)):
        step(x)
    t0 = time.time()
    ms = timed(lambda: step(x), args.steps)
    t1 = time.time()
    value = world * B * args.steps / (ms / 1e3)
    # roofline pass: the same K steps again with every tensor-core convolution launch bracketed by CUDA events
    # on the launching stream (this disables the CUDA-graph replay of the forward, the kernels are identical)
    net.set_profiling(x, True)
    step(x)
    net.profile_collect(x)
    ms_prof = timed(lambda: step(x), args.steps)
    conv_ms, conv_flops, conv_launches = net.profile_collect(x)
    net.set_profiling(x, False)
    # our kernels per step: the score network's (its launch count includes one memset) + what the step call launches
    # (update, scatter, resolve, correct; sdpc_step_kernel_launches) [+ the max merge of the sharded flow]
    step_kernels = run.kernel_launches(p, run.buffers(x, x, x, new_images=new_images)) + (1 if world > 1 else 0)
    launches_per_step = (net.launch_count(x) - 1) + step_kernels

    # ---- end-to-end arm: host buffers, copies inside the timed region ------------------------------
    x_host = g["x"].clone().pin_memory()
    out_host = torch.empty_like(x_host).pin_memory()
    ni_host = torch.empty_like(x_host).pin_memory()
    xd = torch.empty_like(x)

    host = [x_host, out_host]                           # pinned ping-pong: a step's result is the next step's input

    def e2e_step():
        xd.copy_(host[0], non_blocking=True)
        step(xd)
        host[1].copy_(xd, non_blocking=True)
        ni_host.copy_(new_images, non_blocking=True)
        torch.cuda.current_stream().synchronize()       # the caller reads the result before the next step
        host.reverse()

    for _ in range(2):
        e2e_step()
    e2e_steps = max(3, args.steps // 2)
    ms_e2e = timed(e2e_step, e2e_steps)
    e2e_value = world * B * e2e_steps / (ms_e2e / 1e3)
    clk = clocks.stop(t0, t1, time.time()) if clocks else None
    nbytes = x_host.numel() * 4

    # fp32-parity arm on the tensor cores (bf16x3: hi/lo operand split, 2e-4 of the fp32 oracle), same step
    parity = None
    if args.precision == "bf16" and not args.no_parity_arm:
        net3 = NCSN_LiDAR_small(config_ns(dev), precision="bf16x3").to(dev)
        net3.load_state_dict(net.state_dict())
        fast_net, net = net, net3
        for _ in range(3):
            step(x)
        k3 = max(3, args.steps // 2)
        ms3 = timed(lambda: step(x), k3)
        parity = {"dtype": "bf16x3", "value": world * B * k3 / (ms3 / 1e3), "unit": "view-steps/s",
                  "ms_per_step": ms3 / k3, "tolerance": "score within 1e-3 of the fp32 oracle (measured 2e-4, tests/test_gpu_scorenet.py)"}
        net = fast_net
        del net3
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = measured_peaks()
    if args.precision == "bf16":
        peak, peak_src = (peaks["bf16_tflops_sustained"], "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)") \
            if peaks else (1400.0, "fallback 1.4 PFLOP/s sustained (of fallback)")
    else:
        peak, peak_src = ((peaks["bf16_tflops_sustained"] / 2, "half of measured bf16 sustained (tf32 nominal ratio)")
                          if peaks else (700.0, "fallback"))
    achieved = conv_flops / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    line = {
        "metric": "view-steps/sec", "value": value, "unit": "view-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": workload(B, A),
                   "views_per_gpu": B, "group_size": A, "global_views": world * B,
                   "parallelism": f"views x{world} ({B // A} group{'s' if B // A > 1 else ''} of {A} per rank)",
                   "l2": "per-step working set (activations, >2 GB) exceeds the 126 MB L2; no explicit flush",
                   "exchange": "1-float all-reduce(MAX) per step (tooHigh gate)" if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": "view-steps/s", "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": 2 * nbytes,
                "ms_per_step": ms_e2e / e2e_steps},
        "gpu_launches": launches_per_step * args.steps,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "conv_umma_kernel (tcgen05 implicit-GEMM 3x3/1x1 conv)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum per launch, mean of the 6 conv launches captured with
                     # ncu --set full in profiles/r01_ncu_full_conv2.txt (256->256 @32x512x8 views; algorithmic bytes of such
                     # a launch: 67 MB operand + 134 MB fp32 output [+ 134 MB residual])
                     "traffic": 2.18e8 if args.precision == "bf16" else None, "traffic_unit": "bytes/launch (ncu, profiles/)",
                     "peak_source": peak_src, "launches_timed": conv_launches,
                     "conv_share_of_step": conv_ms / ms_prof if ms_prof else None,
                     "timing": "CUDA events around each conv launch, separate eager pass of the same K steps "
                               f"({ms_prof / args.steps:.2f} ms/step without graph replay)",
                     "flops_per_view_forward": net.flops_per_view(x)},
    }
    if parity:
        line["fp32_parity_arm"] = parity
    if world == 1 and not args.no_cpu_baseline:
        v, cms, threads = cpu_view_steps_per_s(3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "view-steps/s", "cores": threads, "kind": "port",
                                "sample": f"3 steps x 1 view (A=1), oracle port (torch CPU ops), {threads} threads, "
                                          f"{cms:.0f} ms/step"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
