#!/bin/bash
# last box call of round 1: GPU test suite, smoke and a short bench with the new defaults (CTA pairs, 128-bit CAS winners)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
( timeout 110 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/f3_tests.log
( timeout 40 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-parity-arm 2>&1 | tail -2 ) > gpurun_out/f3_bench.log
( timeout 30 python __graft_entry__.py --smoke 2>&1 | tail -3 ) > gpurun_out/f3_smoke.log
tail -n 3 gpurun_out/f3_tests.log; cut -c1-400 gpurun_out/f3_bench.log; tail -n 2 gpurun_out/f3_smoke.log
