#!/bin/bash
# bench (both arms) + launch list + full ncu capture of the conv kernel for the final build
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-240; return $rc; }
run smoke python __graft_entry__.py --smoke
run bench_ref python bench.py --impl reference --steps 3 --warmup 1
run bench_bf16 python bench.py
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-arm"
if run bench_plain $CMD; then
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 300 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
  SDPC_NO_GRAPH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 6 -o gpurun_out/prof_conv_final $CMD > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
