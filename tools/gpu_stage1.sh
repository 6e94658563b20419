#!/bin/bash
# first GPU bring-up: staged, each stage in its own process under a timeout
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { name=$1; shift; timeout 600 "$@" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n 25 gpurun_out/$name.log; }
run t_crossview python -m pytest tests/test_gpu_crossview.py -m gpu -q -s -x
run t_score_fp32 python -m pytest tests/test_gpu_scorenet.py -m gpu -q -s -k "fp32"
run t_score_tf32 python -m pytest tests/test_gpu_scorenet.py -m gpu -q -s -k "tf32"
run t_score_bf16 python -m pytest tests/test_gpu_scorenet.py -m gpu -q -s -k "bf16"
run quick_time python tools/quick_time.py 8 tf32,bf16
run quick_time_fp32 python tools/quick_time.py 1 fp32
