#!/bin/bash
# N-GPU weak-scaling check of bench.py (torchrun, NCCL) + 1-GPU reference point
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=${1:-2}
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/scale_1.log 2>&1; echo "n=1 rc=$?"; tail -n 1 gpurun_out/scale_1.log | cut -c1-200
for n in 2 4 8; do
  if [ $n -le $N ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 10 --warmup 3 > gpurun_out/scale_$n.log 2>&1
    echo "n=$n rc=$?"; tail -n 1 gpurun_out/scale_$n.log | cut -c1-200
  fi
done
