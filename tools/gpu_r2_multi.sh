#!/bin/bash
# N GPUs of one box: the NCCL shard test, then bench.py under torchrun (weak scaling + the sharded group)
N=${1:-2}
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -3; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/multi_bench_$N.log 2>&1; echo "bench N=$N rc=$?"
tail -1 gpurun_out/multi_bench_$N.log | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/multi_bench_$N.log').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'sharded', d.get('sharded_group'), 'parity', d.get('fp32_parity_arm',{}).get('value'))
except Exception as e:
    print('parse failed', e); print(open('gpurun_out/multi_bench_$N.log').read()[-3000:])
PY
