#!/bin/bash
# what the driver runs at round end, in its order (GPU tests, smoke, reference arm, bench), then the ncu captures of the
# shipped build that profiles/conv_traffic.json and the launch list are made of
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TO:-900} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-300; return $rc; }
TO=1500 TAILN=3 run f_tests python -m pytest tests -m gpu -q -x
TAILN=3 run f_smoke python -c "import __graft_entry__ as g; g.build(); g.smoke()"
run f_bench_ref python bench.py --impl reference --gpus 1 --steps 20 --warmup 5
SECONDS=0; TO=1500 run f_bench python bench.py --gpus 1 --steps 20 --warmup 5; echo "bench wall seconds: $SECONDS"
for arm in bf16 bf16x3 fp16; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_umma -s 250 -c 8 -o gpurun_out/f_prof_conv_$arm python tools/conv_layers.py 8 $arm 1 > gpurun_out/f_ncu_conv_$arm.log 2>&1; echo "ncu conv $arm rc=$?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"scatter_fast|resolve_kernel|rearm_kernel|fix_winners|langevin_update|correct_kernel" -s 18 -c 6 -o gpurun_out/f_prof_step python tools/time_step.py > gpurun_out/f_ncu_step.log 2>&1; echo "ncu step rc=$?"
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-arm --no-torch-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 320 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
