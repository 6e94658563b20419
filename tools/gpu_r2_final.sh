#!/bin/bash
# what the driver runs at round end, in its order: GPU tests, smoke, reference arm, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TO:-900} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-300; return $rc; }
TO=1500 TAILN=3 run f_tests python -m pytest tests -m gpu -q -x
TAILN=3 run f_smoke python -c "import __graft_entry__ as g; g.build(); g.smoke()"
run f_bench_ref python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
/usr/bin/time -v python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench.log 2> gpurun_out/f_bench.err; echo "bench rc=$?"; tail -1 gpurun_out/f_bench.log | cut -c1-200; grep -E "Elapsed|Maximum resident" gpurun_out/f_bench.err
