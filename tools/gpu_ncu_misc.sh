#!/bin/bash
# full ncu capture of the elementwise kernels of one forward (after a plain run of the same command)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/conv_layers.py 8 bf16 1"
$CMD > gpurun_out/layers_plain.txt 2>&1 || { echo "plain run failed"; tail gpurun_out/layers_plain.txt; exit 1; }
head -1 gpurun_out/layers_plain.txt
timeout 600 ncu --set full --clock-control none -k regex:"to_operand|maxpool5_h2|end_conv_norm|stats_reduce_finalize|begin_conv|meanpool|upsample" -s 150 -c 50 -o gpurun_out/prof_misc $CMD > gpurun_out/ncu_misc.log 2>&1; echo "ncu rc=$?"
