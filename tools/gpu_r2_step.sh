#!/bin/bash
# Round 2: the rebuilt cross-view step (RED-only scatter, per-cell resolve that re-arms, fix pass): tests, timing, ncu.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-300; return $rc; }
TAILN=15 run r2s_tests python -m pytest tests/test_gpu_crossview.py tests/test_zz_gpu_edge_cases.py tests/test_gpu_host_step.py tests/test_gpu_endtoend.py -x -q
STEP="python tools/time_step.py"
for n in 1 2 3 4; do SDPC_XVIEW_BLOCKS_PER_SM=$n $STEP 2>&1 | sed "s/^/blocks_per_sm=$n: /"; done | tee gpurun_out/r2s_blocks_per_sm.log
if run r2s_time_step $STEP; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2s_step_launches.csv $STEP > gpurun_out/r2s_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scatter_fast|resolve_kernel|fix_winners|langevin_update|correct_kernel" -s 15 -c 5 -o gpurun_out/r2s_prof_step $STEP > gpurun_out/r2s_ncu_step.log 2>&1; echo "ncu step rc=$?"
fi
