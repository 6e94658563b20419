#!/bin/bash
# full ncu capture (with source counters) of the swapped-operand conv with operand output: conv launches 55-56 and 66 of the forward
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/conv_layers.py 8 bf16 1"
$CMD > gpurun_out/layers_plain.txt 2>&1 || { echo "plain run failed"; tail gpurun_out/layers_plain.txt; exit 1; }
head -1 gpurun_out/layers_plain.txt
# forwards before the profiled one: 1 (first) + 2 (warm) = 3 x 73 conv launches = 219
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 274 -c 2 -o gpurun_out/prof_swap $CMD > gpurun_out/ncu_swap.log 2>&1; echo "swap rc=$?"
ls -la gpurun_out/*.ncu-rep
