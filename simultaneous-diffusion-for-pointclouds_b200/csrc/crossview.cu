// Langevin update + cross-view consistency step for sm_100a.
//
// Replaces the per-step ATen pipeline of the reference samplers
//   a-4 LiDARGen/models/KITTISampling.py:137-490  (pose matrices)
//   a-5 LiDARGen/models/__init__.py:240-582       (translations)
// (about 25*B small kernels + 3*B radix sorts + 6*B sparse->dense scatters per step) with five launches and no
// memset: update, scatter, resolve, fix (exits at once unless resolve flagged a cell), correct.
//
//   update  : x <- x + eps*g + rho*(-mask*(x-ref)) + s*z ; block max of |x0| -> atomicMax
//   scatter : warp-vote compaction of the source pixels that may contribute; per compacted pixel: decode range (fp32),
//             un-project (fp64), to world, then for every target view of the group: from-world, pixel indices
//             (guarded fp32 estimate, crossview_core.h::pixel_fast), and only for candidates that land in the grid the
//             exact fp64 log-range and five fire-and-forget reductions (RED, nothing is returned to the SM) into the
//             target's R x W grid: min of the key, min of the packed (key | source id), fixed-point depth sum,
//             fixed-point intensity sum, count.  All of them are order independent: two runs are bit-identical.
//   resolve : one thread per grid CELL: average / controlled average, crop + mirror for negative ranges (a cell
//             serves the pixel below it and / or the point-mirrored pixel), existMask, optional newImages - and it
//             re-arms the cell, so the z-buffers are empty again when the call returns.  Where the nearest
//             candidate's identity matters (a "far" cell takes its intensity; debug output) the packed winner is
//             verified by recomputing that source's exact log-range; a cell whose packed winner is not confirmed
//             (needs two candidates within 2e-10 relative, or the test hook) is flagged instead of finished.
//   fix     : only when a cell was flagged: one more traversal of the candidates gives the flagged cells their exact
//             (log-range, smallest source id) winner, the last block to finish completes and re-arms them.
//   correct : x += c * corr with the tooHigh gate.
//
// Compiled with -fmad=false: see crossview_core.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "../../include/sdpc_b200.h"
#include "common.h"
#include "crossview_core.h"

namespace sdpc {

// One z-buffer cell, 32 bytes = one L2 sector.  Armed (empty) state: {~0, ~0, 0, 0} and cnt = 0.
struct __align__(16) CellRec {
  unsigned long long zmin;      // fp64 bits of the smallest squared range (the log-range is a monotone function of it)
  unsigned long long zpack;     // (zmin's bits with the low key_shift bits replaced by the source id); while a cell is
                                // flagged: the smallest source id among the candidates at exactly the nearest range
  long long sum_d;              // fixed-point 2^-40
  long long sum_i;              // fixed-point 2^-32
};
constexpr unsigned kFlagBit = 0x80000000u;          // in cnt: resolve could not confirm the packed winner of this cell
constexpr unsigned long long kMagic = 0x5344504362323030ull;

struct StepHeader {             // first 256 bytes of the workspace
  unsigned int max_bits;        // bit pattern of max |x0| of the last update (sdpc_step_read_max / merge_max use offset 0)
  unsigned int pad0[3];
  unsigned int flag_count;      // cells flagged by the resolve pass of the running call
  unsigned int ticket;          // fix pass: blocks that finished their traversal
  unsigned int pad1[2];
  unsigned long long magic;     // kMagic ^ cells once sdpc_step_workspace_init armed the z-buffers
};

struct StepWorkspace {
  StepHeader* hdr;
  CellRec* rec;                 // [B*R*W]
  unsigned int* cnt;            // [B*R*W]
  unsigned int* work_ctr;       // [2*B] work-item counters of the scatter ([0, n_groups)) and of the fix pass; zero between calls
  int n_groups;
  float* shared_img;            // [B,2,H,W] newImages when the caller does not ask for them
  uint8_t* shared_mask;         // [B,H,W]   imageMask & existMask[0] & sky
  size_t cells;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t workspace_layout(int B, int H, int R, int W, char* base, StepWorkspace* ws) {
  size_t cells = (size_t)B * R * W;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  size_t o_hdr = take(256);
  size_t o_rec = take(cells * sizeof(CellRec));
  size_t o_cnt = take(cells * 4);
  size_t o_ctr = take((size_t)B * 2 * 4);
  size_t o_img = take((size_t)B * 2 * H * W * 4);
  size_t o_msk = take((size_t)B * H * W);
  if (ws) {
    ws->hdr = (StepHeader*)(base + o_hdr);
    ws->rec = (CellRec*)(base + o_rec);
    ws->cnt = (unsigned int*)(base + o_cnt);
    ws->work_ctr = (unsigned int*)(base + o_ctr);
    ws->n_groups = B;             // set by the caller once the group size is known (<= B)
    ws->shared_img = (float*)(base + o_img);
    ws->shared_mask = (uint8_t*)(base + o_msk);
    ws->cells = cells;
  }
  return off;
}

__global__ void __launch_bounds__(256) arm_kernel(StepWorkspace ws) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < ws.cells) {
    CellRec r;
    r.zmin = ~0ull; r.zpack = ~0ull; r.sum_d = 0; r.sum_i = 0;
    ws.rec[i] = r;
    ws.cnt[i] = 0u;
  }
  if (i < 2 * (size_t)ws.n_groups) ws.work_ctr[i] = 0u;
  if (i == 0) {
    ws.hdr->max_bits = 0u;
    ws.hdr->flag_count = 0u;
    ws.hdr->ticket = 0u;
    ws.hdr->magic = kMagic ^ (unsigned long long)ws.cells;
  }
}

// ------------------------------------------------------------------------------------------
// update
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
langevin_update_kernel(float* __restrict__ x, const float* __restrict__ grad, const float* __restrict__ noise,
                       const float* __restrict__ refer, const int32_t* __restrict__ mask,
                       float* __restrict__ gl_out, unsigned int* __restrict__ max_bits,
                       int HW, int v_first, long long n_vec, float eps, float rho, float noise_scale,
                       int do_nan_to_num) {
  // one float4 (4 consecutive pixels of one channel plane) per thread; HW % 4 == 0
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float local_max = 0.0f;
  bool local_nan = false;
  if (i < n_vec) {
    long long e = (long long)v_first * 2 * HW + i * 4;
    float4 xv = *reinterpret_cast<const float4*>(x + e);
    float4 gv = grad ? *reinterpret_cast<const float4*>(grad + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 zv = noise ? *reinterpret_cast<const float4*>(noise + e) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 rv = *reinterpret_cast<const float4*>(refer + e);
    int4 mv = *reinterpret_cast<const int4*>(mask + e);
    if (do_nan_to_num) {
      gv.x = nan_to_num(gv.x); gv.y = nan_to_num(gv.y); gv.z = nan_to_num(gv.z); gv.w = nan_to_num(gv.w);
    }
    float4 gl, o;
    o.x = langevin_value(xv.x, gv.x, rv.x, mv.x, zv.x, eps, rho, noise_scale, &gl.x);
    o.y = langevin_value(xv.y, gv.y, rv.y, mv.y, zv.y, eps, rho, noise_scale, &gl.y);
    o.z = langevin_value(xv.z, gv.z, rv.z, mv.z, zv.z, eps, rho, noise_scale, &gl.z);
    o.w = langevin_value(xv.w, gv.w, rv.w, mv.w, zv.w, eps, rho, noise_scale, &gl.w);
    *reinterpret_cast<float4*>(x + e) = o;
    if (gl_out) *reinterpret_cast<float4*>(gl_out + e) = gl;
    bool is_range_plane = ((e / HW) & 1) == 0;       // channel 0 of its view
    if (is_range_plane) {
      local_max = fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fmaxf(fabsf(o.z), fabsf(o.w)));
      local_nan = (o.x != o.x) || (o.y != o.y) || (o.z != o.z) || (o.w != o.w);
    }
  }
  // torch.max propagates NaN; keep that: a NaN anywhere makes the stored pattern a NaN
  unsigned bits = local_nan ? 0x7fc00000u : __float_as_uint(local_max);
  for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
  __shared__ unsigned warp_max[8];
  if ((threadIdx.x & 31) == 0) warp_max[threadIdx.x >> 5] = bits;
  __syncthreads();
  if (threadIdx.x < 8) {
    bits = warp_max[threadIdx.x];
    for (int o = 4; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffu, bits, o));
    if (threadIdx.x == 0 && bits != 0) atomicMax(max_bits, bits);
  }
}

__global__ void merge_max_kernel(unsigned int* max_bits, const float* other, int n) {
  unsigned b = *max_bits;
  for (int i = 0; i < n; ++i) {
    float v = other[i];
    unsigned ob = (v != v) ? 0x7fc00000u : __float_as_uint(fabsf(v));
    b = max(b, ob);
  }
  *max_bits = b;
}

// ------------------------------------------------------------------------------------------
// view sharding: one exchange per step.  A rank's slot of the gather buffer = its updated x planes followed
// by its max |x0| word; after the all-gather one kernel copies the other ranks' planes into x and folds their maxima in.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
shard_pack_kernel(const float4* __restrict__ x_own, float4* __restrict__ slot, const unsigned int* __restrict__ max_bits,
                  size_t n_vec) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_vec) slot[i] = x_own[i];
  if (i == 0) reinterpret_cast<unsigned int*>(slot + n_vec)[0] = *max_bits;
}

__global__ void __launch_bounds__(256)
shard_unpack_kernel(float4* __restrict__ x, const float4* __restrict__ gbuf, unsigned int* __restrict__ max_bits,
                    int world, int rank, size_t n_vec, size_t slot_vec) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  if (r != rank && i < n_vec) x[(size_t)r * n_vec + i] = gbuf[(size_t)r * slot_vec + i];
  if (r == 0 && i == 0) {              // torch.max over every view of the call (KITTISampling.py:162); a NaN pattern is the largest
    unsigned b = *max_bits;
    for (int k = 0; k < world; ++k) b = max(b, reinterpret_cast<const unsigned int*>(gbuf + (size_t)k * slot_vec + n_vec)[0]);
    *max_bits = b;
  }
}

// ------------------------------------------------------------------------------------------
// scatter / resolve / fix
// ------------------------------------------------------------------------------------------
constexpr int kMaxGroup = 32;
constexpr int kSeg = 128;               // source pixels per work item of the production traversal (one warp, 4 per lane)
constexpr int kTravThreads = 256;       // 8 warps per persistent block, two blocks per SM
constexpr int kTravWarps = kTravThreads / 32;

__device__ const double g_log2_tab[SDPC_LOG2_TABLE_DOUBLES] = {SDPC_LOG2_TABLE};

struct StepArgs {
  float* x;                     // scatter / resolve / fix only read it
  const int32_t* mask;
  const uint8_t* sky;
  const uint8_t* exist;
  const double* to_world;
  const double* from_world;
  const float* origins;
  const double *cos_az, *sin_az, *cos_el, *sin_el;
  float* img;                   // [B,2,H,W] newImages (caller's buffer or workspace scratch)
  int32_t* dbg_row;             // candidate-level debug output: selects the full fp64 scatter
  int32_t* dbg_col;
  uint8_t* dbg_valid;
  int32_t* dbg_cnt;             // cell-level debug output
  int32_t* dbg_winner;
  double* dbg_min_d;
  StepWorkspace ws;
  GeoConsts geo;
  int A, variant, sky_filter, tgt_first, tgt_count;
  float sigma_mod, min_depth_thr;
  double allowance;
  int key_shift;                // low bits of the packed key that hold the source id
  int want_key;                 // the nearest candidate is needed (controlled average or cell-level debug output)
  int winner_mode;              // 0: verify the packed winner where it matters; 1: verify every filled cell;
                                // 2: flag every filled cell (the exact traversal decides every winner)
  int dev_probe;                // SDPC_DEV_HOOKS builds only: SDPC_DEV_PROBE from the environment
};

struct SourcePoint {            // a source pixel in world coordinates (pose variant: homogeneous)
  double wx, wy, wz, ww;
};

// decode + un-project + to-world of pixel p of source view b (a-th view of its group); to: its 4x4 (pose variant)
__device__ __forceinline__ SourcePoint source_point(const StepArgs& a, int b, int src_a, int p, const double* to,
                                                    const float* org) {
  const int HW = a.geo.H * a.geo.W;
  const int r = p / a.geo.W, c = p - r * a.geo.W;
  const float x0 = a.x[((size_t)b * 2) * HW + p];
  const float dist = decode_range(x0, a.sigma_mod, a.geo.recip);
  double P[3];
  unproject(dist, a.cos_az[c], a.sin_az[c], a.cos_el[r], a.sin_el[r], P);
  SourcePoint s;
  s.ww = 1.0;
  if (a.variant == SDPC_VARIANT_POSE) {
    s.wx = dot4(to + 0, P[0], P[1], P[2], 1.0);
    s.wy = dot4(to + 4, P[0], P[1], P[2], 1.0);
    s.wz = dot4(to + 8, P[0], P[1], P[2], 1.0);
    s.ww = dot4(to + 12, P[0], P[1], P[2], 1.0);
  } else {
    s.wx = P[0] + (double)org[src_a * 3 + 0];
    s.wy = P[1] + (double)org[src_a * 3 + 1];
    s.wz = P[2] + (double)org[src_a * 3 + 2];
  }
  return s;
}

// the point in the frame of target view t (from: rows 0..2 of its 4x4; ta: index of t in its group)
__device__ __forceinline__ void to_target(const StepArgs& a, const SourcePoint& s, const double* from, const float* org,
                                          int ta, double* qx, double* qy, double* qz) {
  if (a.variant == SDPC_VARIANT_POSE) {
    *qx = dot4(from + 0, s.wx, s.wy, s.wz, s.ww);
    *qy = dot4(from + 4, s.wx, s.wy, s.wz, s.ww);
    *qz = dot4(from + 8, s.wx, s.wy, s.wz, s.ww);
  } else {
    *qx = s.wx - (double)org[ta * 3 + 0];
    *qy = s.wy - (double)org[ta * 3 + 1];
    *qz = s.wz - (double)org[ta * 3 + 2];
  }
}

// The five reductions of one candidate; none returns a value, so the compiler emits RED and the warp never waits.
// r2: exact squared range (the z-buffer key); nd: its log-range as the depth sum takes it.
__device__ __forceinline__ void accumulate(const StepArgs& a, size_t cell, double r2, double nd, unsigned src_id,
                                           long long inten_fx) {
  CellRec* rec = a.ws.rec + cell;
#ifdef SDPC_DEV_HOOKS                   // timing probe (tools/build_variant.py -D SDPC_DEV_HOOKS, SDPC_DEV_PROBE=1 in the environment):
  if (a.dev_probe == 1) return;         // the traversal without its reductions (a run-time switch keeps the arithmetic alive)
#endif
  if (a.want_key) {
    const unsigned long long key = (unsigned long long)__double_as_longlong(r2);      // r2 >= 0: monotone
    atomicMin(&rec->zmin, key);
    atomicMin(&rec->zpack, ((key >> a.key_shift) << a.key_shift) | (unsigned long long)src_id);
  }
#if defined(SDPC_DEV_PROBE_CNT_IN_REC)    // timing probes (results are wrong): every reduction of a candidate in ONE 32-byte sector
  atomicAdd((unsigned long long*)&rec->sum_d, (unsigned long long)depth_to_fixed_magic(nd));
  atomicAdd((unsigned long long*)&rec->sum_i, (unsigned long long)inten_fx + (1ull << 44));
  return;
#elif defined(SDPC_DEV_PROBE_ONE_RED)     // one 64-bit reduction per candidate
  atomicAdd((unsigned long long*)&rec->sum_d, (unsigned long long)depth_to_fixed_magic(nd));
  return;
#elif defined(SDPC_DEV_PROBE_CNT_ONLY)    // one 32-bit reduction per candidate
  atomicAdd(a.ws.cnt + cell, 1u);
  return;
#endif
  atomicAdd((unsigned long long*)&rec->sum_d, (unsigned long long)depth_to_fixed_magic(nd));
  atomicAdd((unsigned long long*)&rec->sum_i, (unsigned long long)inten_fx);
  atomicAdd(a.ws.cnt + cell, 1u);
}
// fix pass: a candidate of a flagged cell at exactly the nearest range reports its source id
__device__ __forceinline__ void report_winner(const StepArgs& a, size_t cell, double r2, unsigned src_id) {
  if (!(__ldcg(a.ws.cnt + cell) & kFlagBit)) return;
  CellRec* rec = a.ws.rec + cell;
  if ((unsigned long long)__double_as_longlong(r2) == __ldcg(&rec->zmin)) atomicMin(&rec->zpack, (unsigned long long)src_id);
}

// The min-depth filter (KITTISampling.py:273-275, nd > threshold in float64) on the table logarithm: decided by the
// estimate unless it is within 1e-12 of the threshold, where the library expression of the reference decides.
__device__ __forceinline__ bool passes_min_depth(const StepArgs& a, double r2, double nd_fast) {
  if (a.min_depth_thr < 0.0f) return true;
  const double thr = (double)a.min_depth_thr;
  if (fabs(nd_fast - thr) > 1e-12 * (1.0 + thr)) return nd_fast > thr;
  return log_range_of_r2(r2, a.sigma_mod, a.geo) > thr;
}

__device__ __forceinline__ void check_armed(const StepArgs& a) {
  // a workspace that was never armed (sdpc_step_workspace_init) would give silently wrong z-buffers: fail loudly
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0 &&
      a.ws.hdr->magic != (kMagic ^ (unsigned long long)a.ws.cells)) {
    printf("sdpc: step workspace is not armed: call sdpc_step_workspace_init once before the first step\n");
    __trap();
  }
}

// Full fp64 traversal, one thread per source pixel, with the candidate-level debug output: the cross-check of the
// production traversal below (tests assert that both leave bit-identical cells).  FIX: the second traversal that gives
// flagged cells their exact winner.
template <bool FIX>
__device__ __forceinline__ void traverse_full(const StepArgs& a, double* s_to, double* s_from, float* s_org) {
  const int HW = a.geo.H * a.geo.W;
  const int src_a = blockIdx.y, g = blockIdx.z, b = g * a.A + src_a;
  const int t_lo = max(g * a.A, a.tgt_first);
  const int t_hi = min((g + 1) * a.A, a.tgt_first + a.tgt_count);
  if (t_lo >= t_hi) return;
  if (a.variant == SDPC_VARIANT_POSE) {
    if (threadIdx.x < 16) s_to[threadIdx.x] = a.to_world[(size_t)b * 16 + threadIdx.x];
    for (int i = threadIdx.x; i < (t_hi - t_lo) * 12; i += blockDim.x)
      s_from[i] = a.from_world[(size_t)(t_lo + i / 12) * 16 + (i % 12)];
  } else {
    for (int i = threadIdx.x; i < a.A * 3; i += blockDim.x) s_org[i] = a.origins[i];
  }
  __syncthreads();
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= HW) return;
  const bool want_dbg = !FIX && a.dbg_row != nullptr;
  bool src_ok = a.exist[(size_t)src_a * HW + p] != 0;
  if (a.sky_filter) src_ok = src_ok && (a.sky[(size_t)b * HW + p] != 0);
  if (!src_ok && !want_dbg) return;
  const SourcePoint s = source_point(a, b, src_a, p, s_to, s_org);
  const long long inten_fx = FIX ? 0 : inten_to_fixed(a.x[((size_t)b * 2 + 1) * HW + p]);
  const unsigned src_id = (unsigned)(src_a * HW + p);
  const size_t grid_cells = (size_t)a.geo.R * a.geo.W;
  for (int t = t_lo; t < t_hi; ++t) {
    double qx, qy, qz;
    to_target(a, s, s_from + (t - t_lo) * 12, s_org, t - g * a.A, &qx, &qy, &qz);
    const Candidate cd = reproject(qx, qy, qz, a.sigma_mod, a.geo);
    bool ok = src_ok && in_grid(cd, a.geo);
    if (a.min_depth_thr >= 0.0f) ok = ok && (cd.nd > (double)a.min_depth_thr);
    if (want_dbg) {
      const size_t k = (size_t)t * a.A * HW + src_id;
      a.dbg_row[k] = cd.row;
      a.dbg_col[k] = cd.col;
      a.dbg_valid[k] = ok ? 1 : 0;
    }
    if (!ok) continue;
    const size_t cell = (size_t)t * grid_cells + (size_t)cd.row * a.geo.W + cd.col;
    const double r2 = range2(qx, qy, qz);
    if (!FIX) accumulate(a, cell, r2, cd.nd, src_id, inten_fx);
    else report_winner(a, cell, r2, src_id);
  }
}

__global__ void __launch_bounds__(256) scatter_full_kernel(StepArgs a) {
  __shared__ double s_to[16];
  __shared__ double s_from[kMaxGroup * 12];
  __shared__ float s_org[kMaxGroup * 3];
  check_armed(a);
  traverse_full<false>(a, s_to, s_from, s_org);
}

// Production traversal.  Persistent blocks (blockIdx.y = group); every WARP draws work items - 128 consecutive pixels of
// one source view of the group - from the group's counter until none is left, so rows the beam mask empties and
// regions that project outside the target grids do not leave SMs idle.  Per item: warp-vote / prefix compaction of
// the pixels that may contribute (up to 4 per lane, kept in registers as world points), then target by target: the
// target's matrix comes out of shared memory once and serves the lane's pixels.
template <bool FIX>
__device__ __forceinline__ void traverse_items(const StepArgs& a, unsigned* counter) {
  __shared__ double s_to[kMaxGroup * 16];
  __shared__ double s_from[kMaxGroup * 12];
  __shared__ float s_org[kMaxGroup * 3];
  __shared__ double s_tab[SDPC_LOG2_TABLE_DOUBLES];
  __shared__ unsigned char s_list[kTravWarps][kSeg];
  const int HW = a.geo.H * a.geo.W;
  const int g = blockIdx.y;
  const int t_lo = max(g * a.A, a.tgt_first);
  const int t_hi = min((g + 1) * a.A, a.tgt_first + a.tgt_count);
  if (t_lo >= t_hi) return;
  if (a.variant == SDPC_VARIANT_POSE) {
    for (int i = threadIdx.x; i < a.A * 16; i += blockDim.x) s_to[i] = a.to_world[(size_t)g * a.A * 16 + i];
    for (int i = threadIdx.x; i < (t_hi - t_lo) * 12; i += blockDim.x)
      s_from[i] = a.from_world[(size_t)(t_lo + i / 12) * 16 + (i % 12)];
  } else {
    for (int i = threadIdx.x; i < a.A * 3; i += blockDim.x) s_org[i] = a.origins[i];
  }
  for (int i = threadIdx.x; i < SDPC_LOG2_TABLE_DOUBLES; i += blockDim.x) s_tab[i] = g_log2_tab[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned segs = (unsigned)(HW / kSeg), n_items = segs * (unsigned)a.A;
  const unsigned grid_cells = (unsigned)(a.geo.R * a.geo.W);
  unsigned char* list = s_list[warp];
  for (;;) {
    unsigned item = 0;
    if (lane == 0) item = atomicAdd(counter + g, 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= n_items) break;
    const int src_a = (int)(item / segs), base = (int)(item % segs) * kSeg, b = g * a.A + src_a;
    // ---- compaction of the pixels that may contribute
    const uchar4 ex = *reinterpret_cast<const uchar4*>(a.exist + (size_t)src_a * HW + base + lane * 4);
    uchar4 sk = make_uchar4(1, 1, 1, 1);
    if (a.sky_filter) sk = *reinterpret_cast<const uchar4*>(a.sky + (size_t)b * HW + base + lane * 4);
    const unsigned v = (ex.x && sk.x ? 1u : 0u) | (ex.y && sk.y ? 2u : 0u) | (ex.z && sk.z ? 4u : 0u) | (ex.w && sk.w ? 8u : 0u);
    const int mine = __popc(v);
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int n = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += n;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0) continue;
    int off = incl - mine;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (v & (1u << k)) list[off++] = (unsigned char)(lane * 4 + k);
    __syncwarp();
    // ---- the lane's pixels as world points
    SourcePoint sp[4];
    long long ifx[4];
    int pix[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int idx = lane + 32 * j;
      pix[j] = -1;
      ifx[j] = 0;
      sp[j].wx = sp[j].wy = sp[j].wz = 0.0; sp[j].ww = 1.0;
      if (idx < total) {
        const int p = base + list[idx];
        pix[j] = p;
        sp[j] = source_point(a, b, src_a, p, s_to + src_a * 16, s_org);
        if (!FIX) ifx[j] = inten_to_fixed(a.x[((size_t)b * 2 + 1) * HW + p]);
      }
    }
    __syncwarp();                      // the list may be overwritten by the next item from here on
    // ---- target by target: the target's matrix comes out of shared memory once and serves the lane's pixels
    for (int t = t_lo; t < t_hi; ++t) {
      double m[12];
      if (a.variant == SDPC_VARIANT_POSE) {
#pragma unroll
        for (int k = 0; k < 12; ++k) m[k] = s_from[(t - t_lo) * 12 + k];
      } else {
#pragma unroll
        for (int k = 0; k < 3; ++k) m[k] = (double)s_org[(t - g * a.A) * 3 + k];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (pix[j] < 0) continue;
        double qx, qy, qz;
        if (a.variant == SDPC_VARIANT_POSE) {
          qx = dot4(m + 0, sp[j].wx, sp[j].wy, sp[j].wz, sp[j].ww);
          qy = dot4(m + 4, sp[j].wx, sp[j].wy, sp[j].wz, sp[j].ww);
          qz = dot4(m + 8, sp[j].wx, sp[j].wy, sp[j].wz, sp[j].ww);
        } else {
          qx = sp[j].wx - m[0];
          qy = sp[j].wy - m[1];
          qz = sp[j].wz - m[2];
        }
        int row, col;
        if (!pixel_fast(qx, qy, qz, a.geo, &row, &col)) continue;
        const size_t cell = (size_t)((unsigned)t * grid_cells + (unsigned)(row * a.geo.W + col));
        if (FIX) {
          if (!(__ldcg(a.ws.cnt + cell) & kFlagBit)) continue;
        }
        const double r2 = range2(qx, qy, qz);
        const double nd = fast_log_range_of_r2(r2, a.sigma_mod, a.geo, s_tab);
        if (!passes_min_depth(a, r2, nd)) continue;
        const unsigned src_id = (unsigned)(src_a * HW + pix[j]);
        if (!FIX) accumulate(a, cell, r2, nd, src_id, ifx[j]);
        else report_winner(a, cell, r2, src_id);
      }
    }
  }
}

__global__ void __launch_bounds__(kTravThreads) scatter_fast_kernel(StepArgs a) {
  check_armed(a);
  traverse_items<false>(a, a.ws.work_ctr);
}

// Exact squared range of source id `id` of the group of target t, seen from t: the expression the scatter evaluated
__device__ __forceinline__ unsigned long long exact_key_of(const StepArgs& a, int t, unsigned id) {
  const int HW = a.geo.H * a.geo.W;
  const int g = t / a.A, src_a = id / HW, p = id - src_a * HW, b = g * a.A + src_a;
  const SourcePoint s = source_point(a, b, src_a, p, a.to_world + (size_t)b * 16, a.origins);
  double qx, qy, qz;
  to_target(a, s, a.from_world + (size_t)t * 16, a.origins, t - g * a.A, &qx, &qy, &qz);
  return (unsigned long long)__double_as_longlong(range2(qx, qy, qz));
}

// outputs of one cell to the pixel(s) it serves
__device__ __forceinline__ void write_consumers(const StepArgs& a, int t, int gr, int gc, bool use_d, bool use_m,
                                                double depth, float inten, bool filled) {
  const int HW = a.geo.H * a.geo.W, W = a.geo.W, H = a.geo.H, R = a.geo.R;
  if (use_d) {                                                    // the pixel below the crop offset, non-negative range
    const int p = (gr - (R - H)) * W + gc;
    const size_t i0 = ((size_t)t * 2) * HW + p;
    a.img[i0] = (float)depth;
    a.img[i0 + HW] = inten;
    a.ws.shared_mask[(size_t)t * HW + p] = (filled && (a.exist[p] != 0) && (a.sky[(size_t)t * HW + p] != 0)) ? 1 : 0;
  }
  if (use_m) {                                                    // the point-mirrored pixel, negative range
    const int p = (H - 1 - gr) * W + (gc + W / 2 >= W ? gc + W / 2 - W : gc + W / 2);
    const size_t i0 = ((size_t)t * 2) * HW + p;
    a.img[i0] = (float)(depth * -1.0);
    a.img[i0 + HW] = inten;
    a.ws.shared_mask[(size_t)t * HW + p] = (filled && (a.exist[p] != 0) && (a.sky[(size_t)t * HW + p] != 0)) ? 1 : 0;
  }
}

__device__ __forceinline__ void consumers_of(const StepArgs& a, int t, int gr, int gc, bool* use_d, bool* use_m) {
  // crop rows [R-H, R); negative ranges read the point-mirrored cell (KITTISampling.py:401-403): pixel (r, c) reads
  // cell (r + R - H, c) when x0 >= 0 and cell (H - 1 - r, (c - W/2) mod W) when x0 < 0
  const int HW = a.geo.H * a.geo.W, W = a.geo.W, H = a.geo.H, R = a.geo.R;
  const float* x0 = a.x + ((size_t)t * 2) * HW;
  *use_d = (gr >= R - H) && !(x0[(gr - (R - H)) * W + gc] < 0.0f);
  *use_m = (gr < H) && (x0[(H - 1 - gr) * W + (gc + W / 2 >= W ? gc + W / 2 - W : gc + W / 2)] < 0.0f);
}

__device__ __forceinline__ void rearm(const StepArgs& a, size_t cell) {
  CellRec r;
  r.zmin = ~0ull; r.zpack = ~0ull; r.sum_d = 0; r.sum_i = 0;
  a.ws.rec[cell] = r;
  a.ws.cnt[cell] = 0u;
}

// Finish one filled cell whose winner (if anybody needs it) is confirmed: outputs, debug output, re-arm.
__device__ __forceinline__ void finish_cell(const StepArgs& a, int t, int gr, int gc, size_t cell, unsigned cn,
                                            const CellRec& rec, FusedFast f, long long winner) {
  const int HW = a.geo.H * a.geo.W;
  // the nearest log-range: needed by a far cell and by the debug output only (one library log2 + sqrt)
  const double min_d = (a.want_key && (f.far || a.dbg_min_d))
                           ? log_range_of_r2(__longlong_as_double((long long)rec.zmin), a.sigma_mod, a.geo) : 0.0;
  if (f.far) {
    const int g = t / a.A, wa = (int)winner / HW, wp = (int)winner - wa * HW;
    fuse_far(&f, min_d, a.x[((size_t)(g * a.A + wa) * 2 + 1) * HW + wp], a.sigma_mod, a.allowance, a.geo.recip);
  }
  bool use_d, use_m;
  consumers_of(a, t, gr, gc, &use_d, &use_m);
  write_consumers(a, t, gr, gc, use_d, use_m, f.depth, f.inten, true);
  if (a.dbg_cnt) a.dbg_cnt[cell] = (int32_t)cn;
  if (a.dbg_winner) a.dbg_winner[cell] = (int32_t)winner;
  if (a.dbg_min_d) a.dbg_min_d[cell] = min_d;
  rearm(a, cell);
}

// One thread per OUTPUT pixel: the cell it reads (crop, or the point mirror for a negative range), average /
// controlled average, existMask.  Where the nearest candidate's identity matters (a far cell takes its intensity) the
// packed winner is verified first; a cell whose packed winner is not confirmed is flagged and left to the fix pass,
// which then also writes its pixels.
__global__ void __launch_bounds__(256) resolve_kernel(StepArgs a) {
  const int HW = a.geo.H * a.geo.W, W = a.geo.W, H = a.geo.H, R = a.geo.R;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = a.tgt_first + blockIdx.y;
  if (blockIdx.x == 0 && blockIdx.y == 0)                       // the scatter's work counters are free again
    for (int i = threadIdx.x; i < a.ws.n_groups; i += blockDim.x) a.ws.work_ctr[i] = 0u;
  if (p >= HW) return;
  const int r = p / W, c = p - r * W;
  const size_t i0 = ((size_t)t * 2) * HW + p;
  const bool neg = a.x[i0] < 0.0f;
  // crop rows [R-H, R); negative ranges read the point-mirrored cell (KITTISampling.py:401-403)
  const int gr = neg ? (H - 1 - r) : (r + R - H);
  const int gc = neg ? (c - W / 2 < 0 ? c - W / 2 + W : c - W / 2) : c;
  const size_t cell = ((size_t)t * R + gr) * W + gc;
  const unsigned cn = a.ws.cnt[cell] & ~kFlagBit;
  double depth = 0.0;
  float inten = 0.0f;
  if (cn > 0) {
    const CellRec rec = a.ws.rec[cell];
    FusedFast f = fuse_cell_fast(cn, rec.sum_d, rec.sum_i, __longlong_as_double((long long)rec.zmin), a.sigma_mod,
                                 a.allowance, a.geo);
    if (f.far || (a.want_key && a.winner_mode == 2)) {
      const unsigned id = (unsigned)(rec.zpack & ((1ull << a.key_shift) - 1ull));
      const bool confirmed = a.winner_mode != 2 && rec.zpack != ~0ull && exact_key_of(a, t, id) == rec.zmin;
      if (!confirmed) {                                           // leave the cell (and this pixel) to the fix pass
        if (!(atomicOr(a.ws.cnt + cell, kFlagBit) & kFlagBit)) {  // first pixel to flag it (a cell serves up to two)
          a.ws.rec[cell].zpack = ~0ull;
          atomicAdd(&a.ws.hdr->flag_count, 1u);
        }
        return;
      }
      if (f.far) {
        const int g = t / a.A, wa = (int)id / HW, wp = (int)id - wa * HW;
        fuse_far(&f, log_range_of_r2(__longlong_as_double((long long)rec.zmin), a.sigma_mod, a.geo),
                 a.x[((size_t)(g * a.A + wa) * 2 + 1) * HW + wp], a.sigma_mod, a.allowance, a.geo.recip);
      }
    }
    depth = f.depth;
    inten = f.inten;
  }
  // the reference's arithmetic on an empty cell ends in +0 (-0 after the mirror's sign flip) and intensity 0
  a.img[i0] = (float)(neg ? depth * -1.0 : depth);
  a.img[i0 + HW] = inten;
  a.ws.shared_mask[(size_t)t * HW + p] = (cn > 0 && (a.exist[p] != 0) && (a.sky[(size_t)t * HW + p] != 0)) ? 1 : 0;
}

// One thread per grid cell, after resolve: cell-level debug output (with every filled cell's winner verified) and
// re-arming, so that the z-buffers are empty again when the call returns and no step needs a memset.  Flagged cells
// are left to the fix pass.
__global__ void __launch_bounds__(256) rearm_kernel(StepArgs a) {
  const size_t grid_cells = (size_t)a.geo.R * a.geo.W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)a.tgt_count * grid_cells) return;
  const size_t cell = (size_t)a.tgt_first * grid_cells + i;
  const unsigned c = a.ws.cnt[cell];
  if (c == 0) {
    if (a.dbg_cnt) a.dbg_cnt[cell] = 0;
    if (a.dbg_winner) a.dbg_winner[cell] = -1;
    if (a.dbg_min_d) a.dbg_min_d[cell] = 0.0;
    return;
  }
  if (c & kFlagBit) return;
  if (a.want_key && a.winner_mode >= 1) {       // verifying modes (cell-level debug output selects one): every filled cell
    const CellRec rec = a.ws.rec[cell];
    const int t = (int)(cell / grid_cells);
    const unsigned id = (unsigned)(rec.zpack & ((1ull << a.key_shift) - 1ull));
    if (a.winner_mode == 2 || exact_key_of(a, t, id) != rec.zmin) {   // not confirmed: the fix pass finds the exact winner
      a.ws.rec[cell].zpack = ~0ull;
      a.ws.cnt[cell] = c | kFlagBit;
      atomicAdd(&a.ws.hdr->flag_count, 1u);
      return;
    }
    if (a.dbg_cnt) a.dbg_cnt[cell] = (int32_t)c;
    if (a.dbg_winner) a.dbg_winner[cell] = (int32_t)id;
    if (a.dbg_min_d) a.dbg_min_d[cell] = log_range_of_r2(__longlong_as_double((long long)rec.zmin), a.sigma_mod, a.geo);
  }
  rearm(a, cell);
}

// After a fix traversal: the last block to finish completes the flagged cells of this call's target range.
__device__ __forceinline__ void fix_epilogue(const StepArgs& a) {
  __shared__ unsigned s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned total = gridDim.x * gridDim.y * gridDim.z;
    s_last = (atomicAdd(&a.ws.hdr->ticket, 1u) == total - 1u) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const size_t grid_cells = (size_t)a.geo.R * a.geo.W;
  const size_t first = (size_t)a.tgt_first * grid_cells, n = (size_t)a.tgt_count * grid_cells;
  for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
    const size_t cell = first + i;
    const unsigned c = __ldcg(a.ws.cnt + cell);
    if (!(c & kFlagBit)) continue;
    const unsigned cn = c & ~kFlagBit;
    CellRec rec;
    rec.zmin = __ldcg(&a.ws.rec[cell].zmin);
    rec.zpack = __ldcg(&a.ws.rec[cell].zpack);
    rec.sum_d = __ldcg(&a.ws.rec[cell].sum_d);
    rec.sum_i = __ldcg(&a.ws.rec[cell].sum_i);
    const int t = (int)(cell / grid_cells), k = (int)(cell - (size_t)t * grid_cells);
    const FusedFast f = fuse_cell_fast(cn, rec.sum_d, rec.sum_i, __longlong_as_double((long long)rec.zmin), a.sigma_mod,
                                       a.allowance, a.geo);
    long long winner = (long long)rec.zpack;
    if ((unsigned long long)winner >= (unsigned long long)a.A * a.geo.H * a.geo.W) {   // cannot happen: the traversal is
      printf("sdpc: flagged cell %llu found no candidate at its nearest depth\n", (unsigned long long)cell);  // deterministic
      winner = 0;
    }
    finish_cell(a, t, k / a.geo.W, k % a.geo.W, cell, cn, rec, f, winner);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    a.ws.hdr->flag_count = 0u;
    a.ws.hdr->ticket = 0u;
  }
  for (int i = threadIdx.x; i < a.ws.n_groups; i += blockDim.x) a.ws.work_ctr[a.ws.n_groups + i] = 0u;
}

__global__ void __launch_bounds__(kTravThreads) fix_winners_kernel(StepArgs a) {
  if (__ldcg(&a.ws.hdr->flag_count) == 0u) return;                // the usual case: every packed winner was confirmed
  traverse_items<true>(a, a.ws.work_ctr + a.ws.n_groups);
  fix_epilogue(a);
}

__global__ void __launch_bounds__(256) fix_winners_full_kernel(StepArgs a) {
  __shared__ double s_to[16];
  __shared__ double s_from[kMaxGroup * 12];
  __shared__ float s_org[kMaxGroup * 3];
  if (__ldcg(&a.ws.hdr->flag_count) == 0u) return;
  traverse_full<true>(a, s_to, s_from, s_org);
  fix_epilogue(a);
}

// correction (KITTISampling.py:427-430,490): corr = -(imageMask&sky) * (1-mask) * (x - new);
// x += coef * corr, zeroed entirely when the tooHigh gate (KITTISampling.py:162) trips.
__global__ void __launch_bounds__(256)
correct_kernel(float* __restrict__ x, const float* __restrict__ img, const uint8_t* __restrict__ smask,
               const int32_t* __restrict__ mask, const unsigned int* __restrict__ max_bits,
               int32_t* __restrict__ too_high_out, int HW, int v_first, long long n_vec, float sigma_mod,
               float corr_coef, int recip) {
  const float mx = __uint_as_float(*max_bits);
  const bool too_high = too_high_gate(mx, sigma_mod, recip);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && too_high_out) *too_high_out = too_high ? 1 : 0;
  if (i >= n_vec) return;
  long long e = (long long)v_first * 2 * HW + i * 4;
  long long plane = e / HW;                      // view*2 + channel
  long long pix = (plane >> 1) * HW + (e - plane * HW);
  float4 xv = *reinterpret_cast<const float4*>(x + e);
  float4 nv = *reinterpret_cast<const float4*>(img + e);
  int4 mv = *reinterpret_cast<const int4*>(mask + e);
  uchar4 sv = *reinterpret_cast<const uchar4*>(smask + pix);
  float4 o = xv;
  if (!too_high) {
    o.x = xv.x + corr_coef * ((float)(-(int)sv.x * (mv.x == 0 ? 1 : 0)) * (xv.x - nv.x));
    o.y = xv.y + corr_coef * ((float)(-(int)sv.y * (mv.y == 0 ? 1 : 0)) * (xv.y - nv.y));
    o.z = xv.z + corr_coef * ((float)(-(int)sv.z * (mv.z == 0 ? 1 : 0)) * (xv.z - nv.z));
    o.w = xv.w + corr_coef * ((float)(-(int)sv.w * (mv.w == 0 ? 1 : 0)) * (xv.w - nv.w));
  } else {
    o.x = xv.x + corr_coef * 0.0f; o.y = xv.y + corr_coef * 0.0f;
    o.z = xv.z + corr_coef * 0.0f; o.w = xv.w + corr_coef * 0.0f;
  }
  *reinterpret_cast<float4*>(x + e) = o;
}

}  // namespace sdpc

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace sdpc;

static int check_params(const sdpc_step_params* p, const sdpc_step_buffers* b) {
  if (!p || !b) return set_error(SDPC_ERR_ARG, "null params/buffers");
  if (p->n_views <= 0 || p->group_size <= 0 || p->n_views % p->group_size != 0)
    return set_error(SDPC_ERR_ARG, "n_views must be a positive multiple of group_size");
  if (p->group_size > kMaxGroup) return set_error(SDPC_ERR_ARG, "group_size > 32 not supported");
  if (p->height <= 0 || p->width <= 0 || (p->height * p->width) % 4 != 0)
    return set_error(SDPC_ERR_ARG, "H*W must be a positive multiple of 4");
  if ((long long)p->group_size * p->height * p->width > (1ll << 30))
    return set_error(SDPC_ERR_ARG, "group_size*H*W must stay below 2^30 (source ids are packed into the z-buffer key)");
  if (p->tgt_first < 0 || p->tgt_count < 0 || p->tgt_first + p->tgt_count > p->n_views)
    return set_error(SDPC_ERR_ARG, "target range outside [0, n_views)");
  if (!b->x || !b->refer || !b->mask) return set_error(SDPC_ERR_ARG, "x/refer/mask must be non-null");
  return SDPC_OK;
}

// candidate-level debug output selects the full fp64 scatter; so does an image whose H*W is not a multiple of a work item
static bool use_full_scatter(const sdpc_step_params* p, const sdpc_step_buffers* b) {
  const bool dbg_candidates = b->dbg_row && b->dbg_col && b->dbg_valid;
  return dbg_candidates || (p->height * p->width) % kSeg != 0;
}

extern "C" int sdpc_step_kernel_launches(const sdpc_step_params* p, const sdpc_step_buffers* b) {
  if (!p || !b) return set_error(SDPC_ERR_ARG, "null params/buffers");
  return p->share ? 6 : 1;          // update [+ scatter, resolve, re-arm, fix (exits at once unless a cell was flagged), correct]
}

extern "C" size_t sdpc_step_workspace_bytes(int n_views, int height, int width, int big_rows) {
  return workspace_layout(n_views, height, big_rows, width, nullptr, nullptr);
}

extern "C" int sdpc_step_workspace_init(void* workspace, size_t workspace_bytes, int n_views, int height, int width,
                                        int big_rows, void* stream) {
  if (n_views <= 0 || height <= 0 || width <= 0 || big_rows <= 0) return set_error(SDPC_ERR_ARG, "workspace_init: bad shape");
  StepWorkspace ws;
  size_t need = workspace_layout(n_views, height, big_rows, width, (char*)workspace, &ws);
  if (!workspace || workspace_bytes < need) return set_error(SDPC_ERR_WORKSPACE, "step workspace too small");
  arm_kernel<<<(unsigned)((ws.cells + 255) / 256), 256, 0, (cudaStream_t)stream>>>(ws);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_langevin_update(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  if (int e = check_params(p, b)) return e;
  cudaStream_t stream = (cudaStream_t)stream_;
  StepWorkspace ws;
  size_t need = workspace_layout(p->n_views, p->height, p->big_rows, p->width, (char*)workspace, &ws);
  if (!workspace || workspace_bytes < need) return set_error(SDPC_ERR_WORKSPACE, "step workspace too small");
  SDPC_CUDA(cudaMemsetAsync(&ws.hdr->max_bits, 0, sizeof(unsigned), stream));
  const int HW = p->height * p->width;
  const int tcount = p->tgt_count ? p->tgt_count : p->n_views;
  long long n_vec = (long long)tcount * 2 * HW / 4;
  int blocks = (int)((n_vec + 255) / 256);
  langevin_update_kernel<<<blocks, 256, 0, stream>>>(b->x, b->grad, b->noise, b->refer, b->mask, b->grad_likelihood,
                                                     &ws.hdr->max_bits, HW, p->tgt_first, n_vec, p->step_size, p->grad_ref,
                                                     p->noise_scale, p->nan_to_num);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_step_merge_max(void* workspace, const float* other_max, int n, void* stream) {
  if (!workspace || !other_max || n <= 0) return set_error(SDPC_ERR_ARG, "merge_max: bad argument");
  merge_max_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned int*)workspace, other_max, n);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" size_t sdpc_shard_slot_floats(int views_per_rank, int height, int width) {
  return (size_t)views_per_rank * 2 * height * width + 32;       // the planes + one 128-byte line for the max word
}

extern "C" int sdpc_shard_pack(void* workspace, const float* x_own, float* slot, int views_per_rank, int height, int width,
                               void* stream) {
  if (!workspace || !x_own || !slot || views_per_rank <= 0) return set_error(SDPC_ERR_ARG, "shard_pack: bad argument");
  const size_t n = (size_t)views_per_rank * 2 * height * width;
  if (n % 4) return set_error(SDPC_ERR_ARG, "shard_pack: H*W must be a multiple of 2");
  shard_pack_kernel<<<(unsigned)((n / 4 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const float4*)x_own, (float4*)slot, &((StepHeader*)workspace)->max_bits, n / 4);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_shard_unpack(void* workspace, float* x, const float* gathered, int world, int rank, int views_per_rank,
                                 int height, int width, void* stream) {
  if (!workspace || !x || !gathered || world <= 0 || rank < 0 || rank >= world)
    return set_error(SDPC_ERR_ARG, "shard_unpack: bad argument");
  const size_t n = (size_t)views_per_rank * 2 * height * width;
  const size_t slot = sdpc_shard_slot_floats(views_per_rank, height, width);
  const dim3 grid((unsigned)((n / 4 + 255) / 256), world);
  shard_unpack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((float4*)x, (const float4*)gathered,
                                                             &((StepHeader*)workspace)->max_bits, world, rank, n / 4, slot / 4);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_step_read_max(void* workspace, float* out_max, void* stream) {
  if (!workspace || !out_max) return set_error(SDPC_ERR_ARG, "read_max: bad argument");
  SDPC_CUDA(cudaMemcpyAsync(out_max, workspace, sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return SDPC_OK;
}

extern "C" int sdpc_crossview_share(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  if (int e = check_params(p, b)) return e;
  if (!b->sky || !b->exist || !b->cos_az || !b->sin_az || !b->cos_el || !b->sin_el)
    return set_error(SDPC_ERR_ARG, "share: sky/exist/LUT pointers must be non-null");
  if (p->variant == SDPC_VARIANT_POSE && (!b->to_world || !b->from_world))
    return set_error(SDPC_ERR_ARG, "share: pose variant needs to_world/from_world");
  if (p->variant == SDPC_VARIANT_TRANSLATION && !b->origins)
    return set_error(SDPC_ERR_ARG, "share: translation variant needs origins");
  if (p->winner_mode < 0 || p->winner_mode > 2) return set_error(SDPC_ERR_ARG, "winner_mode must be 0, 1 or 2");
  cudaStream_t stream = (cudaStream_t)stream_;
  StepWorkspace ws;
  size_t need = workspace_layout(p->n_views, p->height, p->big_rows, p->width, (char*)workspace, &ws);
  if (!workspace || workspace_bytes < need) return set_error(SDPC_ERR_WORKSPACE, "step workspace too small");
  const int HW = p->height * p->width;
  const int tcount = p->tgt_count ? p->tgt_count : p->n_views;
  const bool dbg_cells = b->dbg_cnt || b->dbg_winner || b->dbg_min_d;

  StepArgs a;
  a.x = b->x; a.mask = b->mask; a.sky = b->sky; a.exist = b->exist;
  a.to_world = b->to_world; a.from_world = b->from_world; a.origins = b->origins;
  a.cos_az = b->cos_az; a.sin_az = b->sin_az; a.cos_el = b->cos_el; a.sin_el = b->sin_el;
  a.img = b->new_images ? b->new_images : ws.shared_img;
  const bool full = use_full_scatter(p, b);
  a.dbg_row = (b->dbg_row && b->dbg_col && b->dbg_valid) ? b->dbg_row : nullptr;
  a.dbg_col = b->dbg_col; a.dbg_valid = b->dbg_valid;
  a.dbg_cnt = b->dbg_cnt; a.dbg_winner = b->dbg_winner; a.dbg_min_d = b->dbg_min_d;
  a.ws = ws;
  a.geo = make_geo(p->h_min, p->dh, p->big_row_min, p->dv, p->height, p->width, p->big_rows, p->scalar_div_recip);
  a.A = p->group_size; a.variant = p->variant; a.sky_filter = p->sky_filter;
  a.tgt_first = p->tgt_first; a.tgt_count = tcount;
  a.sigma_mod = p->sigma_mod; a.min_depth_thr = p->min_depth_thr; a.allowance = p->allowance;
  a.key_shift = 1;
  while ((1 << a.key_shift) < p->group_size * HW) ++a.key_shift;
  if (p->key_shift_override > a.key_shift && p->key_shift_override < 52) a.key_shift = p->key_shift_override;
  a.want_key = (p->allowance >= 0.0 || dbg_cells) ? 1 : 0;
  a.winner_mode = p->winner_mode;
  a.dev_probe = 0;
#ifdef SDPC_DEV_HOOKS
  if (const char* v = getenv("SDPC_DEV_PROBE")) a.dev_probe = atoi(v);
#endif
  if (dbg_cells && a.winner_mode == 0) a.winner_mode = 1;        // the debug output names every cell's winner
  a.ws.n_groups = p->n_views / p->group_size;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // persistent traversal: blocks per group so that the whole grid is about two blocks (16 warps) per SM
  // (SDPC_XVIEW_BLOCKS_PER_SM: tuning knob; measured flat from 2 to 4)
  static const int blocks_per_sm = [] { const char* v = getenv("SDPC_XVIEW_BLOCKS_PER_SM"); int n = v ? atoi(v) : 2; return n > 0 ? n : 2; }();
  const int per_group = std::max(1, std::min((HW / kSeg * p->group_size + kTravWarps - 1) / kTravWarps,
                                             (blocks_per_sm * sms + a.ws.n_groups - 1) / a.ws.n_groups));
  const dim3 sgrid(per_group, a.ws.n_groups);
  if (full) {
    const dim3 grid((HW + 255) / 256, p->group_size, p->n_views / p->group_size);
    scatter_full_kernel<<<grid, 256, 0, stream>>>(a);
  } else {
    scatter_fast_kernel<<<sgrid, kTravThreads, 0, stream>>>(a);
  }
  SDPC_CUDA(cudaGetLastError());
  const dim3 rgrid((unsigned)((HW + 255) / 256), tcount);
  resolve_kernel<<<rgrid, 256, 0, stream>>>(a);
  SDPC_CUDA(cudaGetLastError());
  rearm_kernel<<<(unsigned)(((size_t)tcount * p->big_rows * p->width + 255) / 256), 256, 0, stream>>>(a);
  SDPC_CUDA(cudaGetLastError());
  if (a.want_key) {
    if (HW % kSeg == 0) {
      fix_winners_kernel<<<sgrid, kTravThreads, 0, stream>>>(a);
    } else {
      const dim3 grid((HW + 255) / 256, p->group_size, p->n_views / p->group_size);
      fix_winners_full_kernel<<<grid, 256, 0, stream>>>(a);
    }
    SDPC_CUDA(cudaGetLastError());
  }
  long long n_vec = (long long)tcount * 2 * HW / 4;
  correct_kernel<<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(b->x, a.img, ws.shared_mask, b->mask, &ws.hdr->max_bits,
                                                                     b->too_high, HW, p->tgt_first, n_vec,
                                                                     p->sigma_mod, p->corr_coef, a.geo.recip);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_langevin_reproject_step(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  if (int e = sdpc_langevin_update(p, b, workspace, workspace_bytes, stream)) return e;
  if (p->share) return sdpc_crossview_share(p, b, workspace, workspace_bytes, stream);
  return SDPC_OK;
}

extern "C" int sdpc_langevin_reproject_step_host(const sdpc_step_params* p, const sdpc_step_buffers* b,
                                                 sdpc_score_t* score, const int64_t* labels, void* score_workspace,
                                                 size_t score_workspace_bytes, float* x_host, const float* grad_host,
                                                 const float* noise_host, float* new_images_host, void* workspace,
                                                 size_t workspace_bytes, void* stream_) {
  if (int e = check_params(p, b)) return e;
  if (!x_host) return set_error(SDPC_ERR_ARG, "x_host is null");
  if (score && (!labels || !b->grad)) return set_error(SDPC_ERR_ARG, "score given: labels and b->grad (device) must be non-null");
  if (score && grad_host) return set_error(SDPC_ERR_ARG, "pass either a score handle or grad_host, not both");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t bytes = (size_t)p->n_views * 2 * p->height * p->width * sizeof(float);
  SDPC_CUDA(cudaMemcpyAsync(b->x, x_host, bytes, cudaMemcpyHostToDevice, stream));
  if (score) {
    if (int e = sdpc_score_forward(score, b->x, labels, (float*)b->grad, p->n_views, score_workspace, score_workspace_bytes,
                                   stream_))
      return e;
  } else if (grad_host) {
    if (!b->grad) return set_error(SDPC_ERR_ARG, "grad_host given but b->grad (device staging) is null");
    SDPC_CUDA(cudaMemcpyAsync((void*)b->grad, grad_host, bytes, cudaMemcpyHostToDevice, stream));
  }
  if (noise_host) {
    if (!b->noise) return set_error(SDPC_ERR_ARG, "noise_host given but b->noise (device staging) is null");
    SDPC_CUDA(cudaMemcpyAsync((void*)b->noise, noise_host, bytes, cudaMemcpyHostToDevice, stream));
  }
  if (int e = sdpc_langevin_reproject_step(p, b, workspace, workspace_bytes, stream_)) return e;
  SDPC_CUDA(cudaMemcpyAsync(x_host, b->x, bytes, cudaMemcpyDeviceToHost, stream));
  if (new_images_host && b->new_images && p->share)
    SDPC_CUDA(cudaMemcpyAsync(new_images_host, b->new_images, bytes, cudaMemcpyDeviceToHost, stream));
  return SDPC_OK;
}
