#!/bin/bash
# full ncu capture of the cross-view step kernels (one step), after a plain run of the same command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/time_step.py"
$CMD > gpurun_out/time_step.log 2>&1 || { echo "plain run failed"; tail gpurun_out/time_step.log; exit 1; }
tail -2 gpurun_out/time_step.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scatter_fast|verify_winner|resolve_kernel|langevin_update|correct_kernel" -s 10 -c 5 -o gpurun_out/prof_step $CMD > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"
