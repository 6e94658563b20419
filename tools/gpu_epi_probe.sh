#!/bin/bash
# per-layer conv times with parts of the epilogue dropped (development timing probe; needs a build with SDPC_DEV_HOOKS=1;
# dropped runs give garbage results)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for m in 0 63 62 4 20 1 2 8; do
  SDPC_DEV_EPI_DROP=$m python tools/conv_layers.py 8 bf16 5 > gpurun_out/layers_drop$m.txt 2>&1
  head -1 gpurun_out/layers_drop$m.txt
  sed -n '/by shape/,$p' gpurun_out/layers_drop$m.txt | head -14
done
