#!/bin/bash
# A/B of two builds of the library within one box: forward time, alternating 3 times.  usage: gpu_ab_lib.sh <other.so>
cd "${GRAFT_REPO_ROOT:-/root/repo}"
for i in 1 2 3; do
  echo "--- base"; python tools/quick_time.py 8 bf16 2>&1 | grep forward
  echo "--- $1"; SDPC_LIB=$PWD/$1 python tools/quick_time.py 8 bf16 2>&1 | grep forward
done
