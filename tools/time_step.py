"""Per-kernel view of the Langevin + cross-view step only (run under ncu's launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import sdpc_b200  # noqa
from sdpc_b200 import cabi
from sdpc_b200.step import StepRunner
from tests.golden import cases
B = 8
case = cases.full_multiview(B=B, A=B)
run = StepRunner(case["x"].shape, "cuda:0", case["refer"], case["mask"], case["sky"], case["exist"], B,
                 cabi.SDPC_VARIANT_POSE, to_world=case["toWorld"], from_world=case["fromWorld"])
xx = case["x"].to("cuda:0")
g, z = torch.randn_like(xx), torch.randn_like(xx)
ALLOW = None if os.environ.get("PLAIN_AVERAGE") else 10.0      # PLAIN_AVERAGE=1: no nearest-candidate reductions (3 of 5)
p = run.params(1e-5, 4e-3, 1.0, 0.01, 1.0, True, True, ALLOW, False)
b = run.buffers(xx, g, z)
for _ in range(3):
    run.step(p, b)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    run.step(p, b)
e.record()
torch.cuda.synchronize()
print(f"langevin+crossview step B=A={B}: {a.elapsed_time(e) / 20 * 1e3:.1f} us", flush=True)
