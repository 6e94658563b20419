#!/bin/bash
# Round 2: BASELINE.json configs 3, 4, 5 through bench.py on one GPU, the full-size Line.yml runner, the teacher-forced
# sweeps over the whole schedule, compute-sanitizer on the cross-view tests, and the GPU tests whose bounds were tightened.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout ${TO:-900} "$@" > gpurun_out/$name.log 2>&1; rc=$?; echo "$name rc=$rc"; tail -n ${TAILN:-1} gpurun_out/$name.log | cut -c1-300; return $rc; }
TO=900 TAILN=3 run c_tests python -m pytest tests/test_gpu_scorenet.py tests/test_gpu_endtoend.py tests/test_gpu_teacher_forced.py tests/test_gpu_crossview.py -m gpu -q -s -x
grep -E "^\[|teacher" gpurun_out/c_tests.log | cut -c1-260
for v in inpainting densification; do run c_bench_$v python bench.py --variant $v --steps 20 --warmup 3 --no-cpu-baseline --no-torch-baseline; done
for B in 16 32 64; do run c_bench_views$B python bench.py --views-per-gpu $B --steps 5 --warmup 3 --no-cpu-baseline --no-torch-baseline; done
# teacher-forced sweeps, V = 8, every level
TO=900 run c_tf_bf16x3 python tests/tools/teacher_forced_sweep.py --precision bf16x3 --out gpurun_out/teacher_forced_bf16x3.json
TO=900 run c_tf_bf16 python tests/tools/teacher_forced_sweep.py --precision bf16 --out gpurun_out/teacher_forced_bf16.json
# the full-size runner: Line.yml as shipped (B = 42, A = 7, doThis 0..6, 232 x 5 steps), outputs to /tmp, times to gpurun_out
OUT=/tmp/full_line; rm -rf $OUT
TO=1500 TAILN=2 run c_full_line python -m sdpc_b200.main --sample --ni --config Line.yml --exp $OUT -i out
python - <<'PY' | tee gpurun_out/c_full_line_times.log
import glob, numpy as np, os
for f in sorted(glob.glob('/tmp/full_line/image_samples/out/*_TimeTaken.npy')):
    print(os.path.basename(f).split('_')[0], float(np.load(f)))
for f in sorted(glob.glob('/tmp/full_line/image_samples/out/*_Masked_completion_897.pth.npy')):
    a = np.load(f); print(os.path.basename(f)[:12], a.shape, float(a.min()), float(a.max()), bool(np.isfinite(a).all()))
PY
# compute-sanitizer on the z-buffer atomics
TO=1200 TAILN=6 run c_memcheck compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_crossview.py tests/test_zz_gpu_edge_cases.py -m gpu -q -x -k "not full_size and not samplers"
TO=1200 TAILN=6 run c_racecheck compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_crossview.py -m gpu -q -x -k "bit_exact_vs_device_oracle or rearmed or production_scatter_equals"
