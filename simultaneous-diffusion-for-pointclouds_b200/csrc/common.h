// Error plumbing shared by the translation units of libsdpc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/sdpc_b200.h"

namespace sdpc {
// Records a thread-local message and returns `code` (sdpc_last_error() reads it back).
int set_error(int code, const char* fmt, ...);

// Function attributes (opt-in shared memory, cluster occupancy) belong to a device, not to the process: one-time
// set-up is remembered per device ordinal so that a process driving several GPUs (DataParallel-style) stays correct.
constexpr int kMaxDevices = 64;

// Kernel launch with programmatic dependent launch (PDL): the kernel may start while its predecessor on the stream is
// still draining; every kernel launched this way executes pdl_sync() (griddepcontrol.wait) before its first global
// memory access, so data and buffer-reuse hazards are ordered exactly as without PDL while launch latency and the
// prologue (barrier init, TMEM allocation, weight staging) overlap the predecessor's tail.  Off unless SDPC_PDL=1: on
// this network it measured neutral (implicit trigger) to 2 % slower (early trigger), see DESIGN.md section 4.
template <typename T> struct type_id { typedef T type; };
bool pdl_enabled();
template <typename... P>
static inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                   typename type_id<P>::type... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, args...);
}
#ifdef __CUDACC__
// first statement of every kernel launched through launch_k
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
#ifdef SDPC_PDL_EARLY_TRIGGER
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // measured slower: waiting CTAs crowd the running kernel
#endif
}
#endif

}  // namespace sdpc

#define SDPC_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::sdpc::set_error(SDPC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                             \
  } while (0)
