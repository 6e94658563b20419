// Multi-view dataset assembly: what LiDARGen/datasets/kitti360_im_8Batch.py:94-304 (and its AllForOne /
// densification siblings) do per view around the projection of row N1:
//   (1) move a raw Velodyne scan (float32 x, y, z, remission) into another frame's sensor coordinates:
//       p' = fromWorld . (toWorld . p), two float64 4x4 products per point (:137-141, :178-179);
//   (2) turn the projected depth / remission images into the sampler's inputs (:203-287): holes (depth >= 2057.701 or
//       remission >= 1) join the unknown mask, values are offset by 1e-4, the range goes to log2(d + 1) / 6, both are
//       clipped to [0, 1], the sky mask is shifted down three rows (the three `sky[1:] = sky[:-1]`), and the returned
//       masks are the logical negations.
// Compiled with -fmad=false (numpy rounds every operation).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/sdpc_b200.h"
#include "common.h"

namespace sdpc {

struct Mat4 { double m[16]; };

__device__ __forceinline__ void matvec4(const Mat4& M, const double* v, double* o) {
#pragma unroll
  for (int r = 0; r < 4; ++r)                                  // row . column, k ascending (a 4-term dot product)
    o[r] = ((M.m[r * 4 + 0] * v[0] + M.m[r * 4 + 1] * v[1]) + M.m[r * 4 + 2] * v[2]) + M.m[r * 4 + 3] * v[3];
}

__global__ void __launch_bounds__(256)
transform_scan_kernel(const float* __restrict__ scan, int n, Mat4 to_world, Mat4 from_world, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float4 p = reinterpret_cast<const float4*>(scan)[i];
  const double v[4] = {(double)p.x, (double)p.y, (double)p.z, 1.0};
  double w[4], u[4];
  matvec4(to_world, v, w);
  matvec4(from_world, w, u);
  double* o = out + (size_t)i * 4;
  o[0] = u[0]; o[1] = u[1]; o[2] = u[2];
  o[3] = (double)p.w;                                          // remission rides along untouched (:181)
}

__global__ void __launch_bounds__(256)
postprocess_kernel(const double* __restrict__ depth, const double* __restrict__ intensity,
                   const uint8_t* __restrict__ obf, const uint8_t* __restrict__ sky, int H, int W, double max_range,
                   double* __restrict__ real, uint8_t* __restrict__ known, uint8_t* __restrict__ notsky) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int HW = H * W;
  if (i >= HW) return;
  bool hole = obf ? obf[i] != 0 : false;
  double d = depth[i];
  if (d >= max_range) { hole = true; d = 0.0; }                // :203-204
  d = d + 0.0001;
  d = log2(d + 1.0) / 6.0;                                     // :213
  d = fmin(fmax(d, 0.0), 1.0);                                 // np.clip
  real[i] = d;
  if (intensity) {
    double t = intensity[i];
    if (t >= 1.0) { hole = true; t = 0.0; }                    // :271-272
    t = t + 0.0001;
    t = fmin(fmax(t, 0.0), 1.0);
    real[HW + i] = t;
    if (known) known[HW + i] = hole ? 0 : 1;
  }
  if (known) known[i] = hole ? 0 : 1;                          // np.logical_not(mask), the same mask for both channels
  if (notsky) {
    const int r = i / W, c = i - r * W;
    const int rs = r >= 3 ? r - 3 : 0;                         // three one-row shifts; row 0 is replicated
    notsky[i] = sky[rs * W + c] ? 0 : 1;
  }
}

}  // namespace sdpc

using namespace sdpc;

extern "C" int sdpc_transform_scan(const float* scan, int n_points, const double* to_world, const double* from_world,
                                   double* out, void* stream) {
  if (!scan || !to_world || !from_world || !out) return set_error(SDPC_ERR_ARG, "transform_scan: null argument");
  if (n_points < 0) return set_error(SDPC_ERR_ARG, "transform_scan: negative point count");
  if (n_points == 0) return SDPC_OK;
  Mat4 a, b;
  for (int i = 0; i < 16; ++i) { a.m[i] = to_world[i]; b.m[i] = from_world[i]; }      // host pointers (4x4, row major)
  transform_scan_kernel<<<(n_points + 255) / 256, 256, 0, (cudaStream_t)stream>>>(scan, n_points, a, b, out);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_range_image_postprocess(const double* depth, const double* intensity, const uint8_t* obfuscation,
                                            const uint8_t* sky, int height, int width, double max_range, double* real,
                                            uint8_t* known, uint8_t* notsky, void* stream) {
  if (!depth || !real) return set_error(SDPC_ERR_ARG, "range_image_postprocess: null argument");
  if (notsky && !sky) return set_error(SDPC_ERR_ARG, "range_image_postprocess: notsky output needs the sky input");
  if (height <= 0 || width <= 0) return set_error(SDPC_ERR_ARG, "range_image_postprocess: bad shape");
  const int HW = height * width;
  postprocess_kernel<<<(HW + 255) / 256, 256, 0, (cudaStream_t)stream>>>(depth, intensity, obfuscation, sky, height, width,
                                                                          max_range, real, known, notsky);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}
