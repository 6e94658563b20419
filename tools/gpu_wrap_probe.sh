#!/bin/bash
# dev-build timing probes: epilogue stores / residual loads redirected into a 1 MB window (results are garbage)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
export SDPC_LIB=$PWD/gpurun_ab/lib_dev.so
for w in 0 1 2 3; do
  SDPC_DEV_WRAP=$w python tools/conv_layers.py 8 bf16 5 > gpurun_out/layers_wrap$w.txt 2>&1
  echo "== wrap $w"; head -1 gpurun_out/layers_wrap$w.txt; sed -n "/by shape/,\$p" gpurun_out/layers_wrap$w.txt | head -8
done
SDPC_DEV_EPI_DROP=63 python tools/conv_layers.py 8 bf16 5 > gpurun_out/layers_wrapd.txt 2>&1; echo "== drop 63"; head -1 gpurun_out/layers_wrapd.txt; sed -n "/by shape/,\$p" gpurun_out/layers_wrapd.txt | head -4
