cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_gpu_dist.py -m gpu -q -x 2>&1 | tail -5
