#!/bin/bash
# one group of 8 views on 1 GPU and sharded over N GPUs (all-gather of the updated planes every step)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
python tools/time_sharded_sampler.py 8 8 2>&1 | grep "sharded sampler"
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 tools/time_sharded_sampler.py 8 8 2>&1 | grep "sharded sampler"
  fi
done
