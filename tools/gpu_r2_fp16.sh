#!/bin/bash
# the fp16 arm: tests, smoke, bench with the three arms, ncu of its convolutions, teacher-forced sweep
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TO:-900} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-400; return $rc; }
TO=900 TAILN=2 run h_tests python -m pytest tests/test_gpu_scorenet.py tests/test_gpu_endtoend.py tests/test_gpu_host_step.py -m gpu -q -s -x
grep -E "^\.*\[(fp16|line)" gpurun_out/h_tests.log | cut -c1-300
TAILN=12 run h_smoke python __graft_entry__.py --smoke
grep "smoke:" gpurun_out/h_smoke.log
TO=1200 run h_bench python bench.py --steps 20 --warmup 5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 8 -o gpurun_out/h_prof_conv_fp16 python tools/conv_layers.py 8 fp16 1 > gpurun_out/h_ncu_conv_fp16.log 2>&1; echo "ncu conv fp16 rc=$?"
TO=900 run h_tf_fp16 python tests/tools/teacher_forced_sweep.py --precision fp16 --out gpurun_out/teacher_forced_fp16.json
