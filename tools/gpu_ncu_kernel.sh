#!/bin/bash
# full ncu capture of one kernel by name regex:  gpu_ncu_kernel.sh <regex> <skip> <count> <outname>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/conv_layers.py 8 bf16 1"
$CMD > gpurun_out/layers_plain.txt 2>&1 || { echo "plain run failed"; tail gpurun_out/layers_plain.txt; exit 1; }
head -1 gpurun_out/layers_plain.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -o gpurun_out/$4 $CMD > gpurun_out/ncu_$4.log 2>&1; echo "ncu rc=$?"
