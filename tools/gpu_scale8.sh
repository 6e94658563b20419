#!/bin/bash
# 8-GPU (or $1-GPU) weak-scaling point of bench.py plus the 2-GPU NCCL parity tests
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-8}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/scale_$N.log 2>&1
echo "n=$N rc=$?"; tail -n 1 gpurun_out/scale_$N.log | cut -c1-400
timeout 600 python -m pytest tests/test_gpu_dist.py -q -m gpu 2>&1 | tail -2
