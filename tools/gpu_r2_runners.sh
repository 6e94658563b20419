#!/bin/bash
# the other two runners at full size (B = 42, A = 7, 232 x 5 steps): Inpainting.yml (7 arms) and Densification.yml (2 arms), bf16;
# one Line.yml arm in the fp32-parity arm for the time table
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for cfg in Inpainting Densification; do
  OUT=/tmp/full_$cfg; rm -rf $OUT
  timeout 1500 python -m sdpc_b200.main --sample --ni --config $cfg.yml --exp $OUT -i out > gpurun_out/r_full_$cfg.log 2>&1; echo "$cfg rc=$?"
  python - <<PY | tee gpurun_out/r_full_${cfg}_times.log
import glob, numpy as np, os
for f in sorted(glob.glob('/tmp/full_$cfg/image_samples/out/*_TimeTaken.npy')):
    print('$cfg', os.path.basename(f).split('_')[0], float(np.load(f)))
for f in sorted(glob.glob('/tmp/full_$cfg/image_samples/out/*_Masked_completion_897.pth.npy')):
    a = np.load(f); print('$cfg', os.path.basename(f)[:12], a.shape, float(a.min()), float(a.max()), bool(np.isfinite(a).all()))
print('$cfg', 'shared files', len(glob.glob('/tmp/full_$cfg/image_samples/out/*_Shared_completion_initial897.pth.npy')))
PY
  rm -rf $OUT
done
