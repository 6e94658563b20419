#!/bin/bash
# First box call of round 2 (about 10 box-minutes): the evidence the last session of round 1 could not collect.
#   1. full test suite, smoke, both bench arms (plain runs)
#   2. ncu launch list of the bench command (per-kernel share of the step)
#   3. ncu --set full of the CTA-pair convolution (CG = 2) and, for comparison, the multicast variant (SDPC_CTA2=0)
#   4. ncu --set full of the cross-view step kernels with the 128-bit CAS winners
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-300; return $rc; }
TAILN=3 run r2_tests python -m pytest tests -m gpu -x -q
run r2_smoke python __graft_entry__.py --smoke
run r2_bench_ref python bench.py --impl reference --steps 3 --warmup 1
run r2_bench python bench.py
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-arm"
if run r2_bench_plain $CMD; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 300 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
fi
LAY="python tools/conv_layers.py 8 bf16 1"
if run r2_layers $LAY; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 6 -o gpurun_out/r2_prof_conv_cg2 $LAY > gpurun_out/r2_ncu_conv_cg2.log 2>&1; echo "ncu conv cg2 rc=$?"
  SDPC_CTA2=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 6 -o gpurun_out/r2_prof_conv_mc $LAY > gpurun_out/r2_ncu_conv_mc.log 2>&1; echo "ncu conv multicast rc=$?"
fi
STEP="python tools/time_step.py"
if run r2_time_step $STEP; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scatter_fast|resolve_kernel|langevin_update|correct_kernel" -s 8 -c 4 -o gpurun_out/r2_prof_step $STEP > gpurun_out/r2_ncu_step.log 2>&1; echo "ncu step rc=$?"
fi
# bf16x3 arm in clusters / CTA pairs: only when tools/patches/x3_cluster.patch has been applied and built (it adds the
# SDPC_X3_CLUSTER switch); bit-for-bit against the single-CTA path, then the forward time of both
if grep -q conv_umma_x3_cluster simultaneous-diffusion-for-pointclouds_b200/csrc/conv_umma.h; then
  timeout 120 python tools/ab_probe.py run gpurun_out/ab_def 0 > gpurun_out/r2_ab_def.log 2>&1
  SDPC_X3_CLUSTER=1 timeout 120 python tools/ab_probe.py run gpurun_out/ab_x3cl 0 > gpurun_out/r2_ab_x3cl.log 2>&1
  python tools/ab_probe.py compare gpurun_out/ab_def gpurun_out/ab_x3cl | tee gpurun_out/r2_ab_x3_cmp.log
  python tools/quick_time.py 8 bf16x3 2>&1 | grep forward | tee gpurun_out/r2_x3_time.log
  SDPC_X3_CLUSTER=1 python tools/quick_time.py 8 bf16x3 2>&1 | grep forward | tee -a gpurun_out/r2_x3_time.log
  rm -f gpurun_out/ab_*.npy
fi
for B in 1 2 4; do python tools/quick_time.py $B bf16 2>&1 | grep forward; done > gpurun_out/r2_small_batches.log
cat gpurun_out/r2_small_batches.log
# BASELINE config 5 on one GPU: 16 / 32 / 64 views in groups of 8 (forward + step per call)
for B in 16 32 64; do python tools/quick_time.py $B bf16 8 2>&1 | grep -E "forward|step"; done > gpurun_out/r2_view_sweep.log
cat gpurun_out/r2_view_sweep.log
# the same switch in a variant library (tools/build_variant.py gpurun_ab/lib_x3cl.so tools/patches/x3_cluster.patch): the switch is
# off by default, so the generic loop below checks that the variant equals the shipped build; this block turns it on
if [ -f gpurun_ab/lib_x3cl.so ]; then
  timeout 120 python tools/ab_probe.py run gpurun_out/ab_def 0 > gpurun_out/r2_ab_def.log 2>&1
  SDPC_LIB=$PWD/gpurun_ab/lib_x3cl.so SDPC_X3_CLUSTER=1 timeout 120 python tools/ab_probe.py run gpurun_out/ab_x3cl_on 0 > gpurun_out/r2_ab_x3cl_on.log 2>&1
  python tools/ab_probe.py compare gpurun_out/ab_def gpurun_out/ab_x3cl_on | tee gpurun_out/r2_ab_x3cl_on_cmp.log
  for i in 1 2; do
    python tools/quick_time.py 8 bf16x3 2>&1 | grep forward
    SDPC_LIB=$PWD/gpurun_ab/lib_x3cl.so SDPC_X3_CLUSTER=1 python tools/quick_time.py 8 bf16x3 2>&1 | grep forward
  done | tee gpurun_out/r2_x3cl_time.log
fi
# library variants built from tools/patches with tools/build_variant.py (gpurun_ab/*.so travels with the snapshot, the shipped
# library is untouched): score-network outputs bit for bit against the shipped build, then the forward time of both, alternating
if ls gpurun_ab/*.so > /dev/null 2>&1; then
  timeout 120 python tools/ab_probe.py run gpurun_out/ab_def 0 > gpurun_out/r2_ab_def.log 2>&1
  for lib in gpurun_ab/*.so; do
    name=$(basename $lib .so)
    SDPC_LIB=$PWD/$lib timeout 120 python tools/ab_probe.py run gpurun_out/ab_$name 0 > gpurun_out/r2_ab_$name.log 2>&1
    python tools/ab_probe.py compare gpurun_out/ab_def gpurun_out/ab_$name | tee gpurun_out/r2_ab_${name}_cmp.log
    bash tools/gpu_ab_lib.sh $lib 2>&1 | tee gpurun_out/r2_ab_${name}_time.log
  done
  rm -f gpurun_out/ab_*.npy
fi
# the same sweep through bench.py (full JSON line per view count: value, e2e, roofline)
for B in 16 32 64; do timeout 600 python bench.py --views-per-gpu $B --steps 5 --warmup 3 --no-cpu-baseline --no-parity-arm 2>&1 | tail -1; done > gpurun_out/r2_bench_sweep.log
cut -c1-260 gpurun_out/r2_bench_sweep.log
