cd "${GRAFT_REPO_ROOT:-/root/repo}"
python -m pytest tests/test_zz_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -8
