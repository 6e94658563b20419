#!/bin/bash
# full ncu captures of the main kernels (one bench invocation per capture, short)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity-arm"
export SDPC_NO_GRAPH=1
$CMD > gpurun_out/bench_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/bench_plain.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on -k regex:maxpool5 -s 10 -c 2 -o gpurun_out/prof_maxpool $CMD > gpurun_out/ncu_maxpool.log 2>&1; echo "maxpool rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 6 -o gpurun_out/prof_conv2 $CMD > gpurun_out/ncu_conv2.log 2>&1; echo "conv rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:to_operand -s 40 -c 2 -o gpurun_out/prof_toop $CMD > gpurun_out/ncu_toop.log 2>&1; echo "toop rc=$?"
ls -la gpurun_out/*.ncu-rep
