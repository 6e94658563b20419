#!/bin/bash
# parity tests + smoke + bench + ncu launch list + one full ncu capture of the conv kernel
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-12} gpurun_out/$name.log; return $rc; }
TAILN=30 run t_gpu python -m pytest tests -m gpu -q -s
run smoke python __graft_entry__.py --smoke
run bench_bf16 python bench.py --steps 10 --warmup 3
run bench_tf32 python bench.py --steps 10 --warmup 3 --precision tf32 --no-cpu-baseline --no-parity-arm
run bench_ref python bench.py --impl reference --steps 2 --warmup 1
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-arm"
if run bench_plain $CMD; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 900 -c 360 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 120 -c 3 -o gpurun_out/prof_conv $CMD > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
fi
ls -la gpurun_out
