#!/bin/bash
# ncu captures of the shipped build (conv kernels of the three 16-bit arms, the step kernels, the launch list of bench.py);
# the reports are summarised ON the box (they exceed what gpurun copies back) and removed
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for arm in bf16 bf16x3 fp16; do
  timeout 900 ncu --set full --clock-control none -k regex:conv_umma -s 250 -c 8 -o /tmp/f_prof_conv_$arm python tools/conv_layers.py 8 $arm 1 > gpurun_out/f_ncu_conv_$arm.log 2>&1; echo "ncu conv $arm rc=$?"
  python tools/ncu_summary.py /tmp/f_prof_conv_$arm.ncu-rep > gpurun_out/f_ncu_conv_$arm.txt
done
python tools/ncu_traffic.py bf16=/tmp/f_prof_conv_bf16.ncu-rep bf16x3=/tmp/f_prof_conv_bf16x3.ncu-rep fp16=/tmp/f_prof_conv_fp16.ncu-rep --source "profiles/r02_ncu_conv_<arm>.txt (ncu --set full --clock-control none of tools/conv_layers.py 8 <arm> 1, 8 launches from the 250th of the forward; tools/gpu_r2_ncu.sh)" --out gpurun_out/f_conv_traffic.json
timeout 600 ncu --set full --clock-control none -k regex:"scatter_fast|resolve_kernel|rearm_kernel|fix_winners|langevin_update|correct_kernel" -s 18 -c 6 -o /tmp/f_prof_step python tools/time_step.py > gpurun_out/f_ncu_step.log 2>&1; echo "ncu step rc=$?"
python tools/ncu_summary.py /tmp/f_prof_step.ncu-rep > gpurun_out/f_ncu_step.txt
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-parity-arm --no-torch-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 800 -c 320 --csv --log-file gpurun_out/f_launches.csv $CMD > gpurun_out/f_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/f_bench.log 2>&1; tail -1 gpurun_out/f_bench.log | cut -c1-200
du -sh gpurun_out
