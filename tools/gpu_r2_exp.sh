cd "${GRAFT_REPO_ROOT:-/root/repo}"
S="python tools/time_step.py"
L=$PWD/gpurun_ab/lib_hooks.so
for v in "SDPC_LIB=$L" "SDPC_LIB=$L SDPC_DEV_PROBE=1" "SDPC_LIB=$L SDPC_DEV_PROBE=1 SDPC_XVIEW_BLOCKS_PER_SM=4" "SDPC_LIB=$L SDPC_DEV_PROBE=1 SDPC_XVIEW_BLOCKS_PER_SM=1"; do
 env $v ncu --metrics gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv -k regex:"scatter_fast" -s 5 -c 1 $S 2>&1 | grep -E "scatter_fast" | awk -F'","' '{print "'"$v"' " $5 " " $(NF-2) " " $NF}' | cut -c1-220
done
