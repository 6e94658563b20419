// Point cloud -> range image: the dataloader-side projection that produces the sampler's
// inputs.  Replaces LiDARGen/datasets/lidar_utils.py:54-347 (numpy argsort + np.unique(axis=1) + scipy coo_matrix
// per scan) with the same z-buffer scatter family as the cross-view step:
//   scatter<0> : per point spherical projection (float64, numpy's operation order), clamp to the image, in-grid test,
//                atomicMin of the float64 depth bits per pixel
//   scatter<1> : the point whose depth equals the pixel minimum claims it (atomicMin on the point index: ties go to
//                the smallest index; the reference breaks them arbitrarily)
//   resolve    : depth / planar range / remission / point-index images, flipped by 180 degrees as np.flip does
//   obfuscation: the row-sequential scan of lidar_utils.py:262-305, one thread per column
// Compiled with -fmad=false (numpy evaluates one rounded operation at a time).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/sdpc_b200.h"
#include "common.h"

namespace sdpc {

constexpr double kMaxRange = 2057.701;

struct ProjGeo {
  double ox, oy, oz, h_min, dh, v_min, dv;
  int H, W;
};

struct ProjPoint {
  double depth, xy;
  int row, col;
  bool ok;
};

__device__ __forceinline__ int clamp_round(double v, int hi) {
  const double r = rint(v);                                  // np.round: half to even; astype(int) then clip
  if (!(r >= 0.0)) return 0;                                 // negatives and NaN clamp to 0 like np.maximum(0, .)
  return r > (double)hi ? hi : (int)r;
}

__device__ __forceinline__ ProjPoint project_point(const double* __restrict__ pt, const ProjGeo& g) {
  ProjPoint p;
  const double x = pt[0] - g.ox, y = pt[1] - g.oy, z = pt[2] - g.oz;
  const double xy2 = x * x + y * y;
  p.depth = sqrt(xy2 + z * z);
  const double horiz = atan2(y, x);
  p.xy = sqrt(xy2);
  const double vert = atan2(z, p.xy);
  p.col = clamp_round((horiz - g.h_min) / g.dh, g.W - 1);
  p.row = clamp_round((vert - g.v_min) / g.dv, g.H - 1);
  p.ok = p.col > 0 && p.row > 0;                             // lidar_utils.py:166: strictly positive after the clamp
  return p;
}

template <int PASS>
__global__ void __launch_bounds__(256)
proj_scatter_kernel(const double* __restrict__ pts, int n, int stride, ProjGeo g, unsigned long long* __restrict__ zmin,
                    unsigned int* __restrict__ winner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const ProjPoint p = project_point(pts + (size_t)i * stride, g);
  if (!p.ok || !(p.depth == p.depth)) return;
  const int pix = p.row * g.W + p.col;
  const unsigned long long key = (unsigned long long)__double_as_longlong(p.depth);
  if (PASS == 0) atomicMin(zmin + pix, key);
  else if (zmin[pix] == key) atomicMin(winner + pix, (unsigned)i);
}

__global__ void __launch_bounds__(256)
proj_resolve_kernel(const double* __restrict__ pts, int stride, int intensity_col, ProjGeo g,
                    const unsigned int* __restrict__ winner, double* __restrict__ depth, double* __restrict__ xy,
                    double* __restrict__ inten, double* __restrict__ index) {
  const int o = blockIdx.x * blockDim.x + threadIdx.x;       // output (flipped) pixel
  const int HW = g.H * g.W;
  if (o >= HW) return;
  const int src = HW - 1 - o;                                // np.flip over both axes
  const unsigned w = winner[src];
  double d = kMaxRange, r = kMaxRange, it = 0.0, id = -1.0;
  if (w != 0xFFFFFFFFu) {
    const ProjPoint p = project_point(pts + (size_t)w * stride, g);
    if (p.depth != 0.0) {                                    // a nearest depth of exactly 0 reads as "empty" (:183)
      d = p.depth;
      r = p.xy;
      id = (double)w;
      if (intensity_col >= 0) it = pts[(size_t)w * stride + intensity_col];
    }
  }
  depth[o] = d;
  xy[o] = r;
  index[o] = id;
  if (inten) inten[o] = it;
}

// lidar_utils.py:262-305.  One block; thread c owns column c (W <= 1024): its running minimum and the sky flag of the
// previous row live in registers, the three-row "differs from the minimum" counts are exchanged through shared memory.
__global__ void __launch_bounds__(1024)
proj_obfuscation_kernel(const double* __restrict__ xy, uint8_t* __restrict__ obf, uint8_t* __restrict__ sky, int H, int W) {
  __shared__ int e[1024 + 2];
  const int c = threadIdx.x;
  const bool act = c < W;
  double min_depth = kMaxRange;
  bool sky_prev = true;
  if (c == 0) { e[0] = 0; e[W + 1] = 0; }
  if (act) {
    obf[c] = 0;
    obf[W + c] = 0;
    for (int r = 0; r < H; ++r) sky[r * W + c] = 0;            // the reference clears the sky mask before returning
  }
  for (int r = 2; r < H - 1; ++r) {
    double cur = 0.0;
    if (act) {
      cur = xy[r * W + c];
      obf[r * W + c] = cur > min_depth + 5.0 ? 1 : 0;
      e[c + 1] = (cur != min_depth) + (xy[(r - 1) * W + c] != min_depth) + (xy[(r + 1) * W + c] != min_depth);
    }
    __syncthreads();
    if (act) {
      const bool eq = (e[c + 1] + e[c] + e[c + 2]) <= 1;
      const bool cur_sky = eq && sky_prev;
      sky_prev = cur_sky;
      if (!cur_sky) min_depth = fmin(cur, min_depth);
    }
    __syncthreads();
  }
  if (act && H >= 3) obf[(H - 1) * W + c] = xy[(H - 1) * W + c] > min_depth + 5.0 ? 1 : 0;
}

}  // namespace sdpc

using namespace sdpc;

extern "C" size_t sdpc_projection_workspace_bytes(int height, int width) {
  const size_t hw = (size_t)height * width;
  return hw * 8 /*zmin*/ + hw * 4 /*winner*/ + hw * 8 /*planar range*/ + 1024;
}

extern "C" int sdpc_pointcloud_to_range_image(const sdpc_projection_params* p, const double* points, double* depth,
                                              double* intensity, uint8_t* obfuscation, uint8_t* sky, double* index,
                                              void* workspace, size_t workspace_bytes, void* stream_) {
  if (!p || !points || !depth || !obfuscation || !sky || !index) return set_error(SDPC_ERR_ARG, "projection: null argument");
  if (p->n_points < 0 || p->point_stride < 3 || p->height < 3 || p->width < 1 || p->width > 1024)
    return set_error(SDPC_ERR_ARG, "projection: need stride >= 3, H >= 3, 1 <= W <= 1024");
  if (p->intensity_col >= p->point_stride) return set_error(SDPC_ERR_ARG, "projection: intensity column outside the point");
  if (p->intensity_col >= 0 && !intensity) return set_error(SDPC_ERR_ARG, "projection: intensity output is null");
  const size_t hw = (size_t)p->height * p->width;
  if (!workspace || workspace_bytes < sdpc_projection_workspace_bytes(p->height, p->width))
    return set_error(SDPC_ERR_WORKSPACE, "projection workspace too small");
  cudaStream_t stream = (cudaStream_t)stream_;
  char* base = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
  unsigned long long* zmin = (unsigned long long*)base;
  double* xy = (double*)(base + hw * 8);
  unsigned int* winner = (unsigned int*)(base + hw * 16);
  ProjGeo g;
  g.ox = p->origin[0]; g.oy = p->origin[1]; g.oz = p->origin[2];
  g.h_min = p->h_min; g.dh = p->dh; g.v_min = p->v_min; g.dv = p->dv;
  g.H = p->height; g.W = p->width;
  SDPC_CUDA(cudaMemsetAsync(zmin, 0xFF, hw * 8, stream));
  SDPC_CUDA(cudaMemsetAsync(winner, 0xFF, hw * 4, stream));
  if (p->n_points > 0) {
    const int blocks = (p->n_points + 255) / 256;
    proj_scatter_kernel<0><<<blocks, 256, 0, stream>>>(points, p->n_points, p->point_stride, g, zmin, winner);
    proj_scatter_kernel<1><<<blocks, 256, 0, stream>>>(points, p->n_points, p->point_stride, g, zmin, winner);
    SDPC_CUDA(cudaGetLastError());
  }
  proj_resolve_kernel<<<(unsigned)((hw + 255) / 256), 256, 0, stream>>>(points, p->point_stride, p->intensity_col, g, winner, depth,
                                                                       xy, p->intensity_col >= 0 ? intensity : nullptr, index);
  SDPC_CUDA(cudaGetLastError());
  proj_obfuscation_kernel<<<1, 1024, 0, stream>>>(xy, obfuscation, sky, p->height, p->width);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}
