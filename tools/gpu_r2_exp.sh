cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_gpu_scorenet.py -m gpu -q -s -x -k "fp16" 2>&1 | grep -E "fp16|passed|failed" | cut -c1-500
