cd "${GRAFT_REPO_ROOT:-/root/repo}"
for i in 1 2; do
for B in 1 2; do
  python tools/quick_time.py $B bf16 2>&1 | grep forward | sed 's/^/default     : /'
  SDPC_CTA2=0 python tools/quick_time.py $B bf16 2>&1 | grep forward | sed 's/^/CTA2=0      : /'
  SDPC_CLUSTER=0 python tools/quick_time.py $B bf16 2>&1 | grep forward | sed 's/^/CLUSTER=0   : /'
  SDPC_NO_GRAPH=1 python tools/quick_time.py $B bf16 2>&1 | grep forward | sed 's/^/NO_GRAPH=1  : /'
done; done
