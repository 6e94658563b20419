cd "${GRAFT_REPO_ROOT:-/root/repo}"
L=$PWD/gpurun_ab/lib_ilp4.so
SDPC_LIB=$L python -m pytest tests/test_gpu_crossview.py tests/test_zz_gpu_edge_cases.py -m gpu -q -x 2>&1 | tail -4
for i in 1 2; do python tools/time_step.py | sed 's/^/base: /'; SDPC_LIB=$L python tools/time_step.py | sed 's/^/ilp4: /'; for n in 1 3; do SDPC_XVIEW_BLOCKS_PER_SM=$n SDPC_LIB=$L python tools/time_step.py | sed "s/^/ilp4 blocks=$n: /"; done; done
SDPC_LIB=$L ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:"scatter_fast|resolve_kernel" -s 10 -c 2 python tools/time_step.py 2>&1 | grep -E "scatter_fast|resolve_kernel" | awk -F'","' '{print $5 " " $NF}'
