#!/bin/bash
# One short box call: (1) cross-view parity suite with the 128-bit CAS winner path, step time of both paths,
# (2) score network with CTA pairs (SDPC_CTA2=1) against the shipped clusters: bit-for-bit outputs and forward time.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( SDPC_XVIEW_CAS128=1 timeout 120 python -m pytest tests/test_gpu_crossview.py -x -q 2>&1 | tail -12 ) > gpurun_out/ab_xview_tests.log
( timeout 60 python tools/time_step.py; SDPC_XVIEW_CAS128=1 timeout 60 python tools/time_step.py ) > gpurun_out/ab_xview_time.log 2>&1
timeout 80 python tools/ab_probe.py run gpurun_out/ab_def 8 > gpurun_out/ab_def.log 2>&1
SDPC_CTA2=1 timeout 80 python tools/ab_probe.py run gpurun_out/ab_cta2 8 > gpurun_out/ab_cta2.log 2>&1
python tools/ab_probe.py compare gpurun_out/ab_def gpurun_out/ab_cta2 > gpurun_out/ab_cmp.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,power.draw --format=csv > gpurun_out/ab_smi.log 2>&1
tail -4 gpurun_out/ab_xview_tests.log gpurun_out/ab_xview_time.log gpurun_out/ab_def.log gpurun_out/ab_cta2.log gpurun_out/ab_cmp.log
rm -f gpurun_out/ab_*.npy
