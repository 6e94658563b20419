#!/bin/bash
# Round 2 main evidence run: full GPU suite with the measured errors printed, smoke, bench (both arms + torch-GPU baseline),
# step timing, ncu of the shipped convolution kernels and of the step kernels.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TO:-900} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-400; return $rc; }
TO=1500 TAILN=3 run m_tests python -m pytest tests -m gpu -q -s -x
grep -E "^\[|rel err|flips|teacher" gpurun_out/m_tests.log | cut -c1-260
TAILN=4 run m_smoke python __graft_entry__.py --smoke
run m_bench_ref python bench.py --impl reference --steps 8 --warmup 2
TO=1500 run m_bench python bench.py
python tools/time_step.py | tee gpurun_out/m_time_step.log
for B in 1 2; do python tools/quick_time.py $B bf16 2>&1 | grep forward; SDPC_PDL=1 python tools/quick_time.py $B bf16 2>&1 | grep forward | sed 's/^/PDL=1 /'; done | tee gpurun_out/m_small_batches.log
python tools/conv_layers.py 1 bf16 3 > gpurun_out/m_layers_B1.log 2>&1; head -3 gpurun_out/m_layers_B1.log
LAY="python tools/conv_layers.py 8 bf16 1"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 8 -o gpurun_out/m_prof_conv_bf16 $LAY > gpurun_out/m_ncu_conv_bf16.log 2>&1; echo "ncu conv bf16 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 6 -o gpurun_out/m_prof_conv_x3 python tools/conv_layers.py 8 bf16x3 1 > gpurun_out/m_ncu_conv_x3.log 2>&1; echo "ncu conv x3 rc=$?"
STEP="python tools/time_step.py"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scatter_fast|resolve_kernel|rearm_kernel|fix_winners|langevin_update|correct_kernel" -s 18 -c 6 -o gpurun_out/m_prof_step $STEP > gpurun_out/m_ncu_step.log 2>&1; echo "ncu step rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-arm --no-torch-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 320 --csv --log-file gpurun_out/m_launches.csv $CMD > gpurun_out/m_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
