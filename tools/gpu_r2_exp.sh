cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/g_bench.log 2>&1; python - <<'PY'
import json
l=json.loads(open('gpurun_out/g_bench.log').read().strip().splitlines()[-1])
print('bf16', round(l['value'],1), 'e2e', round(l['e2e']['value'],1), 'frac', round(l['roofline']['frac'],3), 'traffic', l['roofline']['traffic'])
for k in ('tf32_class_arm','fp32_parity_arm'):
    a=l[k]; print(k, round(a['value'],1), round(a['e2e']['value'],1), round(a['roofline']['frac'],3))
print(l['clocks'])
PY
