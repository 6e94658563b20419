// Error plumbing shared by the translation units of libsdpc_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "../../include/sdpc_b200.h"

namespace sdpc {
// Records a thread-local message and returns `code` (sdpc_last_error() reads it back).
int set_error(int code, const char* fmt, ...);
}  // namespace sdpc

#define SDPC_CUDA(expr)                                                                         \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess)                                                                      \
      return ::sdpc::set_error(SDPC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                               __FILE__, __LINE__);                                             \
  } while (0)
