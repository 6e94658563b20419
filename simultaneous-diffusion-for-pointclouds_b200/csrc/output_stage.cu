// Output stage of the sampler: what runs immediately after the Langevin loop.
//   (1) range image -> xyz point cloud, LiDARGen/visualization.py:12-43: depth = 2^(6 r) - 1 (float32), the fixed
//       yaw / pitch grid of the 64 x 1024 sensor, points kept iff 0.5 < depth < 63, in row-major pixel order;
//   (2) the error sums of MeasureResults/QuantifyingNotebookSynthesis_Line.ipynb (cell 1): per view, the L1 depth and
//       intensity errors against the ground truth over all pixels and over the input (known) pixels.
// (1) is a stream compaction: every block counts the valid pixels of its 1024-pixel chunk, the chunk offsets are the
// exclusive scan of those counts, and a second pass writes the surviving points with a warp-vote / popc prefix, so
// the output order is the reference's boolean-mask order.  (2) is a fixed-tree reduction in float64.
// Compiled with -fmad=false: numpy rounds every product of cos(yaw) * cos(pitch) * depth separately.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/sdpc_b200.h"
#include "common.h"

namespace sdpc {

constexpr int kChunk = 1024;                    // pixels per block (= threads per block)

__device__ __forceinline__ float unlog_depth(float r) { return exp2f(r * 6.0f) - 1.0f; }   // np.exp2(r*6)-1 in float32

// count[v][chunk] = number of pixels of the chunk with 0.5 < depth < 63
__global__ void __launch_bounds__(kChunk)
points_count_kernel(const float* __restrict__ image, int HW, int chunks, int* __restrict__ count) {
  const int v = blockIdx.y, c = blockIdx.x, p = c * kChunk + threadIdx.x;
  bool ok = false;
  if (p < HW) {
    const float d = unlog_depth(image[(size_t)v * 2 * HW + p]);
    ok = d > 0.5f && d < 63.0f;
  }
  const int n = __syncthreads_count(ok);
  if (threadIdx.x == 0) count[v * chunks + c] = n;
}

// xyz[v][k] for the k-th valid pixel of view v (row-major order), intensity[v][k], pixel[v][k]; total[v] = #valid
__global__ void __launch_bounds__(kChunk)
points_write_kernel(const float* __restrict__ image, int HW, int W, int chunks, const int* __restrict__ count,
                    const double* __restrict__ cos_yaw, const double* __restrict__ sin_yaw,
                    const double* __restrict__ cos_pitch, const double* __restrict__ sin_pitch,
                    double* __restrict__ xyz, float* __restrict__ intensity, int* __restrict__ pixel,
                    int* __restrict__ total) {
  __shared__ int warp_base[kChunk / 32];
  __shared__ int chunk_base;
  const int v = blockIdx.y, c = blockIdx.x, p = c * kChunk + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {                                     // exclusive scan of the view's chunk counts (<= 64 values)
    int b = 0;
    for (int i = 0; i < c; ++i) b += count[v * chunks + i];
    chunk_base = b;
    if (c == chunks - 1 && total) total[v] = b + count[v * chunks + c];
  }
  float d = 0.0f;
  bool ok = false;
  if (p < HW) {
    d = unlog_depth(image[(size_t)v * 2 * HW + p]);
    ok = d > 0.5f && d < 63.0f;
  }
  const unsigned ballot = __ballot_sync(0xffffffffu, ok);
  if (lane == 0) warp_base[warp] = __popc(ballot);
  __syncthreads();
  if (threadIdx.x == 0) {
    int b = 0;
    for (int i = 0; i < kChunk / 32; ++i) { const int t = warp_base[i]; warp_base[i] = b; b += t; }
  }
  __syncthreads();
  if (!ok) return;
  const int k = chunk_base + warp_base[warp] + __popc(ballot & ((1u << lane) - 1u));
  const int row = p / W, col = p - row * W;
  const double dd = (double)d;
  const size_t o = (size_t)v * HW + k;
  xyz[o * 3 + 0] = cos_yaw[col] * cos_pitch[row] * dd;         // np.cos(yaw) * np.cos(pitch) * depth, left to right
  xyz[o * 3 + 1] = -sin_yaw[col] * cos_pitch[row] * dd;
  xyz[o * 3 + 2] = sin_pitch[row] * dd;
  if (intensity) intensity[o] = image[(size_t)v * 2 * HW + HW + p];
  if (pixel) pixel[o] = p;
}

// out[v][8] = { sum|dp-dg| all, sum|ip-ig| all, sum|dp-dg| input, sum|ip-ig| input, sum dp input, #all, #input, 0 }
// with d = 2^(6 r) - 1 (float32), input = (input_r > 0.001) & (dg < 63); one block per view, fixed reduction tree.
__global__ void __launch_bounds__(1024)
error_sums_kernel(const float* __restrict__ pred, const float* __restrict__ gt, const float* __restrict__ input, int HW,
                  double* __restrict__ out) {
  __shared__ double red[32][7];
  const int v = blockIdx.x;
  const float *pr = pred + (size_t)v * 2 * HW, *gr = gt + (size_t)v * 2 * HW, *ir = input + (size_t)v * 2 * HW;
  double s[7] = {0, 0, 0, 0, 0, 0, 0};
  for (int p = threadIdx.x; p < HW; p += blockDim.x) {
    const float dp = unlog_depth(pr[p]), dg = unlog_depth(gr[p]);
    const float ed = fabsf(dp - dg), ei = fabsf(pr[HW + p] - gr[HW + p]);
    s[0] += (double)ed;
    s[1] += (double)ei;
    s[5] += 1.0;
    if (ir[p] > 0.001f && dg < 63.0f) {
      s[2] += (double)ed;
      s[3] += (double)ei;
      s[4] += (double)dp;
      s[6] += 1.0;
    }
  }
#pragma unroll
  for (int k = 0; k < 7; ++k)
    for (int o = 16; o > 0; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
  if ((threadIdx.x & 31) == 0)
    for (int k = 0; k < 7; ++k) red[threadIdx.x >> 5][k] = s[k];
  __syncthreads();
  if (threadIdx.x < 7) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w][threadIdx.x];
    out[v * 8 + threadIdx.x] = t;
  }
  if (threadIdx.x == 7) out[v * 8 + 7] = 0.0;
}

}  // namespace sdpc

using namespace sdpc;

extern "C" size_t sdpc_points_workspace_bytes(int n_views, int height, int width) {
  if (n_views <= 0 || height <= 0 || width <= 0) return 0;
  const int chunks = (height * width + kChunk - 1) / kChunk;
  return (size_t)n_views * chunks * sizeof(int) + 256;
}

extern "C" int sdpc_range_image_to_points(const float* image, int n_views, int height, int width, const double* cos_yaw,
                                          const double* sin_yaw, const double* cos_pitch, const double* sin_pitch,
                                          double* xyz, float* intensity, int* pixel, int* n_points, void* workspace,
                                          size_t workspace_bytes, void* stream) {
  if (!image || !cos_yaw || !sin_yaw || !cos_pitch || !sin_pitch || !xyz || !n_points || !workspace)
    return set_error(SDPC_ERR_ARG, "range_image_to_points: null argument");
  if (n_views <= 0 || height <= 0 || width <= 0) return set_error(SDPC_ERR_ARG, "range_image_to_points: bad shape");
  if (workspace_bytes < sdpc_points_workspace_bytes(n_views, height, width))
    return set_error(SDPC_ERR_WORKSPACE, "range_image_to_points: workspace too small");
  const int HW = height * width, chunks = (HW + kChunk - 1) / kChunk;
  if (chunks > 65535 || n_views > 65535) return set_error(SDPC_ERR_UNSUPPORTED, "range_image_to_points: image too large");
  cudaStream_t s = (cudaStream_t)stream;
  int* count = (int*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  dim3 grid(chunks, n_views);
  points_count_kernel<<<grid, kChunk, 0, s>>>(image, HW, chunks, count);
  SDPC_CUDA(cudaGetLastError());
  points_write_kernel<<<grid, kChunk, 0, s>>>(image, HW, width, chunks, count, cos_yaw, sin_yaw, cos_pitch, sin_pitch, xyz,
                                              intensity, pixel, n_points);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_depth_intensity_errors(const float* pred, const float* gt, const float* input, int n_views, int height,
                                           int width, double* out, void* stream) {
  if (!pred || !gt || !input || !out) return set_error(SDPC_ERR_ARG, "depth_intensity_errors: null argument");
  if (n_views <= 0 || height <= 0 || width <= 0) return set_error(SDPC_ERR_ARG, "depth_intensity_errors: bad shape");
  error_sums_kernel<<<n_views, 1024, 0, (cudaStream_t)stream>>>(pred, gt, input, height * width, out);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}
