#!/bin/bash
# ncu launch list (gpu__time_duration) of ~2 steps of the bench, after a plain run of the same command
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-arm"
export SDPC_NO_GRAPH=1
$CMD > gpurun_out/bench_plain.log 2>&1 || { echo "plain run failed"; tail gpurun_out/bench_plain.log; exit 1; }
tail -1 gpurun_out/bench_plain.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${1:-800} -c ${2:-330} --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1; echo "ncu rc=$?"
