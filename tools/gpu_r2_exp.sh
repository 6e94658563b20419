cd "${GRAFT_REPO_ROOT:-/root/repo}"
for i in 1 2; do
  python tools/quick_time.py 8 bf16 2>&1 | grep forward | sed 's/^/base: /'
  for lib in gpurun_ab/lib_op_*.so; do SDPC_LIB=$PWD/$lib python tools/quick_time.py 8 bf16 2>&1 | grep forward | sed "s#^#$(basename $lib .so): #"; done
done
