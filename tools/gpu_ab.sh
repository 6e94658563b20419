#!/bin/bash
# A/B of an env-switchable variant within one box: forward time, alternating 3 times
cd "${GRAFT_REPO_ROOT:-/root/repo}"
VAR=$1
for i in 1 2 3; do
  echo "--- base"; python tools/quick_time.py 8 bf16 2>&1 | grep forward
  echo "--- $VAR=1"; env $VAR=1 python tools/quick_time.py 8 bf16 2>&1 | grep forward
done
