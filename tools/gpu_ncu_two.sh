#!/bin/bash
# full ncu captures (source counters) of two conv launches of the 4th forward: indices $1 and $2 of the 73 convs
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python tools/conv_layers.py 8 bf16 1"
$CMD > gpurun_out/layers_plain.txt 2>&1 || { echo "plain run failed"; tail gpurun_out/layers_plain.txt; exit 1; }
head -1 gpurun_out/layers_plain.txt
for idx in $1 $2; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s $((219 + idx)) -c 1 -o gpurun_out/prof_conv_i$idx $CMD > gpurun_out/ncu_i$idx.log 2>&1; echo "ncu idx $idx rc=$?"
done
