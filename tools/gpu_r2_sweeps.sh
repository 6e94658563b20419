#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TO:-900} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-600; return $rc; }
TO=900 run c_tf_bf16x3 python tests/tools/teacher_forced_sweep.py --precision bf16x3 --out gpurun_out/teacher_forced_bf16x3.json
TO=900 run c_tf_bf16 python tests/tools/teacher_forced_sweep.py --precision bf16 --out gpurun_out/teacher_forced_bf16.json
TO=600 TAILN=3 run c_tests2 python -m pytest tests/test_gpu_teacher_forced.py -m gpu -q -s -x
