// Langevin update + cross-view consistency step for sm_100a.
//
// Replaces the per-step ATen pipeline of the reference samplers
//   a-4 LiDARGen/models/KITTISampling.py:137-490  (pose matrices)
//   a-5 LiDARGen/models/__init__.py:240-582       (translations)
// (about 25*B small kernels + 3*B radix sorts + 6*B sparse->dense scatters per step) with four
// launches: update, scatter (z-buffer build incl. the nearest candidate), resolve, correct.
//
//   update  : x <- x + eps*g + rho*(-mask*(x-ref)) + s*z ; block max of |x0| -> atomicMax
//   scatter : one thread per SOURCE pixel: decode range (fp32), un-project (fp64), to world,
//             then for every target view of the group: from-world, spherical re-projection,
//             validity, and order-independent atomics into the target's R x W grid:
//             atomicAdd count / fixed-point depth sum / fixed-point intensity sum, and the nearest
//             candidate as {fp64 bits of the log-range, source id}: lexicographic minimum by a
//             128-bit compare-and-swap (deterministic tie break: smallest source id).
//   winner  : legacy / cross-check paths only (candidate-level debug output, winner_mode = 1): 64-bit
//             atomicMin on the key, then the candidate whose log-range equals the grid minimum claims
//             the pixel, either through a packed key + verification or a second traversal.
//   resolve : one thread per OUTPUT pixel: average / controlled average, crop + mirror for
//             negative ranges, existMask, correction, in-place x update, optional newImages.
//
// Compiled with -fmad=false: see crossview_core.h.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/sdpc_b200.h"
#include "common.h"
#include "crossview_core.h"

namespace sdpc {

struct StepWorkspace {
  unsigned int* max_bits;       // [1]   bit pattern of max |x0|
  unsigned long long* zmin;     // [B*R*W] fp64 bits of the nearest log-range (0xFF.. = empty)
  unsigned int* winner;         // [B*R*W] source id of the nearest candidate
  long long* sum_d;             // [B*R*W] fixed-point 2^-40
  long long* sum_i;             // [B*R*W] fixed-point 2^-32
  unsigned int* cnt;            // [B*R*W]
  unsigned long long* zpack;    // [B*R*W] (log-range bits with the low key_shift bits replaced by the source id)
  unsigned int* flag;           // [1]   set when a packed winner could not be confirmed -> exact winner pass runs
  ulonglong2* zkey;             // [B*R*W] default winner path: {.x = source id, .y = fp64 bits of its log-range}, the
                                //         lexicographic minimum over (.y, .x) kept by a 128-bit compare-and-swap
  float* shared_img;            // [B,2,H,W] newImages when the caller does not ask for them
  uint8_t* shared_mask;         // [B,H,W]   imageMask & existMask[0] & sky
  size_t cells;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static size_t workspace_layout(int B, int H, int R, int W, char* base, StepWorkspace* ws) {
  size_t cells = (size_t)B * R * W;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
  size_t o_max = take(256);
  size_t o_zmin = take(cells * 8);
  size_t o_win = take(cells * 4);
  size_t o_sd = take(cells * 8);
  size_t o_si = take(cells * 8);
  size_t o_cnt = take(cells * 4);
  size_t o_zpack = take(cells * 8);
  size_t o_zkey = take(cells * 16);
  size_t o_img = take((size_t)B * 2 * H * W * 4);
  size_t o_msk = take((size_t)B * H * W);
  if (ws) {
    ws->shared_img = (float*)(base + o_img);
    ws->shared_mask = (uint8_t*)(base + o_msk);
    ws->max_bits = (unsigned int*)(base + o_max);
    ws->zmin = (unsigned long long*)(base + o_zmin);
    ws->winner = (unsigned int*)(base + o_win);
    ws->sum_d = (long long*)(base + o_sd);
    ws->sum_i = (long long*)(base + o_si);
    ws->cnt = (unsigned int*)(base + o_cnt);
    ws->zpack = (unsigned long long*)(base + o_zpack);
    ws->zkey = (ulonglong2*)(base + o_zkey);
    ws->flag = (unsigned int*)(base + o_max + 64);
    ws->cells = cells;
  }
  return off;
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
// scatter / winner
// ------------------------------------------------------------------------------------------
constexpr int kMaxGroup = 32;

struct ScatterArgs {
  const float* x;
  const uint8_t* sky;
  const uint8_t* exist;
  const double* to_world;
  const double* from_world;
  const float* origins;
  const double *cos_az, *sin_az, *cos_el, *sin_el;
  int32_t* dbg_row;
  int32_t* dbg_col;
  uint8_t* dbg_valid;
  StepWorkspace ws;
  GeoConsts geo;
  int A, variant, sky_filter, tgt_first, tgt_count;
  float sigma_mod, min_depth_thr;
  int key_shift;                // low bits of the packed key that hold the source id
};

// Exact (log-range, source id) minimum in ONE 16-byte word per cell (the default winner path): the nearest candidate and,
// among candidates at exactly the same depth, the smallest source id - the same winner the packed-key path confirms
// with its verification pass, without that pass and without the second 64-bit atomicMin.  Values only ever decrease,
// so a stale (even torn) first read can only cost one extra CAS round, never a wrong skip.
__device__ __forceinline__ ulonglong2 cas128(ulonglong2* p, ulonglong2 cmp, ulonglong2 val) {
  ulonglong2 old;
  asm volatile(
      "{\n\t.reg .b128 c, s, o;\n\t"
      "mov.b128 c, {%2, %3};\n\t"
      "mov.b128 s, {%4, %5};\n\t"
      "atom.relaxed.gpu.global.cas.b128 o, [%6], c, s;\n\t"
      "mov.b128 {%0, %1}, o;\n\t}"
      : "=l"(old.x), "=l"(old.y)
      : "l"(cmp.x), "l"(cmp.y), "l"(val.x), "l"(val.y), "l"(p)
      : "memory");
  return old;
}
__device__ __forceinline__ void zkey_min(ulonglong2* p, unsigned long long key, unsigned src_id) {
  const ulonglong2 mine = make_ulonglong2((unsigned long long)src_id, key);
  ulonglong2 cur = *p;
  while (mine.y < cur.y || (mine.y == cur.y && mine.x < cur.x)) {
    const ulonglong2 old = cas128(p, cur, mine);
    if (old.x == cur.x && old.y == cur.y) break;
    cur = old;
  }
}

template <int PASS>
__global__ void __launch_bounds__(256) scatter_kernel(ScatterArgs a) {
  const int HW = a.geo.H * a.geo.W;
  const int src_a = blockIdx.y;              // source view within its group
  const int g = blockIdx.z;                  // group
  const int b = g * a.A + src_a;             // global source view
  // targets of this group that this call resolves
  const int t_lo = max(g * a.A, a.tgt_first);
  const int t_hi = min((g + 1) * a.A, a.tgt_first + a.tgt_count);
  if (t_lo >= t_hi) return;
  if (PASS == 1 && *a.ws.flag == 0) return;   // every packed winner was confirmed: nothing to do

  __shared__ double s_to[16];
  __shared__ double s_from[kMaxGroup * 12];
  __shared__ float s_org[kMaxGroup * 3];
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
  const bool want_dbg = (PASS == 0) && (a.dbg_row != nullptr);
  bool src_ok = a.exist[(size_t)src_a * HW + p] != 0;
  if (a.sky_filter) src_ok = src_ok && (a.sky[(size_t)b * HW + p] != 0);
  if (!src_ok && !want_dbg) return;

  const int r = p / a.geo.W, c = p - r * a.geo.W;
  const float x0 = a.x[((size_t)b * 2) * HW + p];
  const float x1 = a.x[((size_t)b * 2 + 1) * HW + p];
  const float dist = decode_range(x0, a.sigma_mod, a.geo.recip);
  double P[3];
  unproject(dist, a.cos_az[c], a.sin_az[c], a.cos_el[r], a.sin_el[r], P);
  double wx, wy, wz, ww = 1.0;
  if (a.variant == SDPC_VARIANT_POSE) {
    wx = dot4(s_to + 0, P[0], P[1], P[2], 1.0);
    wy = dot4(s_to + 4, P[0], P[1], P[2], 1.0);
    wz = dot4(s_to + 8, P[0], P[1], P[2], 1.0);
    ww = dot4(s_to + 12, P[0], P[1], P[2], 1.0);
  } else {
    wx = P[0] + (double)s_org[src_a * 3 + 0];
    wy = P[1] + (double)s_org[src_a * 3 + 1];
    wz = P[2] + (double)s_org[src_a * 3 + 2];
  }
  const long long inten_fx = inten_to_fixed(x1);
  const unsigned src_id = (unsigned)(src_a * HW + p);
  const size_t grid_cells = (size_t)a.geo.R * a.geo.W;

  for (int t = t_lo; t < t_hi; ++t) {
    double qx, qy, qz;
    if (a.variant == SDPC_VARIANT_POSE) {
      const double* m = s_from + (t - t_lo) * 12;
      qx = dot4(m + 0, wx, wy, wz, ww);
      qy = dot4(m + 4, wx, wy, wz, ww);
      qz = dot4(m + 8, wx, wy, wz, ww);
    } else {
      const int ta = t - g * a.A;
      qx = wx - (double)s_org[ta * 3 + 0];
      qy = wy - (double)s_org[ta * 3 + 1];
      qz = wz - (double)s_org[ta * 3 + 2];
    }
    Candidate cd = reproject(qx, qy, qz, a.sigma_mod, a.geo);
    bool ok = src_ok && in_grid(cd, a.geo);
    if (a.min_depth_thr >= 0.0f) ok = ok && (cd.nd > (double)a.min_depth_thr);
    if (want_dbg) {
      size_t k = (size_t)t * a.A * HW + src_id;
      a.dbg_row[k] = cd.row;
      a.dbg_col[k] = cd.col;
      a.dbg_valid[k] = ok ? 1 : 0;
    }
    if (!ok) continue;
    const size_t cell = (size_t)t * grid_cells + (size_t)cd.row * a.geo.W + cd.col;
    const unsigned long long key = (unsigned long long)__double_as_longlong(cd.nd);   // nd >= 0: monotone
    if (PASS == 0) {
      atomicMin(a.ws.zmin + cell, key);
      atomicMin(a.ws.zpack + cell, ((key >> a.key_shift) << a.key_shift) | (unsigned long long)src_id);
      atomicAdd(a.ws.cnt + cell, 1u);
      atomicAdd((unsigned long long*)(a.ws.sum_d + cell), (unsigned long long)depth_to_fixed(cd.nd));
      atomicAdd((unsigned long long*)(a.ws.sum_i + cell), (unsigned long long)inten_fx);
    } else {
      if (a.ws.zmin[cell] == key) atomicMin(a.ws.winner + cell, src_id);
    }
  }
}


// ------------------------------------------------------------------------------------------
// production scatter: compacted valid source pixels, fp32-guarded re-projection; nearest candidate by a 128-bit CAS on
// {log-range, source id} (CAS = true, default) or by the packed 64-bit key that verify_winner_kernel confirms (CAS = false)
// ------------------------------------------------------------------------------------------
constexpr int kChunk = 1024;      // source pixels per block (4 per thread)

template <bool CAS>
__global__ void __launch_bounds__(256) scatter_fast_kernel(ScatterArgs a) {
  const int HW = a.geo.H * a.geo.W;
  const int src_a = blockIdx.y, g = blockIdx.z, b = g * a.A + src_a;
  const int t_lo = max(g * a.A, a.tgt_first);
  const int t_hi = min((g + 1) * a.A, a.tgt_first + a.tgt_count);
  if (t_lo >= t_hi) return;
  __shared__ double s_to[16];
  __shared__ double s_from[kMaxGroup * 12];
  __shared__ float s_org[kMaxGroup * 3];
  __shared__ unsigned short s_list[kChunk];
  __shared__ int s_wsum[8];
  if (a.variant == SDPC_VARIANT_POSE) {
    if (threadIdx.x < 16) s_to[threadIdx.x] = a.to_world[(size_t)b * 16 + threadIdx.x];
    for (int i = threadIdx.x; i < (t_hi - t_lo) * 12; i += blockDim.x)
      s_from[i] = a.from_world[(size_t)(t_lo + i / 12) * 16 + (i % 12)];
  } else {
    for (int i = threadIdx.x; i < a.A * 3; i += blockDim.x) s_org[i] = a.origins[i];
  }
  // ---- phase 1: warp-vote / prefix compaction of the source pixels that may contribute
  const int base = blockIdx.x * kChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uchar4 ex = *reinterpret_cast<const uchar4*>(a.exist + (size_t)src_a * HW + base + threadIdx.x * 4);
  uchar4 sk = make_uchar4(1, 1, 1, 1);
  if (a.sky_filter) sk = *reinterpret_cast<const uchar4*>(a.sky + (size_t)b * HW + base + threadIdx.x * 4);
  const unsigned v = (ex.x && sk.x ? 1u : 0u) | (ex.y && sk.y ? 2u : 0u) | (ex.z && sk.z ? 4u : 0u) | (ex.w && sk.w ? 8u : 0u);
  const int mine = __popc(v);
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) s_wsum[warp] = incl;
  __syncthreads();
  int off = incl - mine, total = 0;
#pragma unroll
  for (int w = 0; w < 8; ++w) {
    if (w < warp) off += s_wsum[w];
    total += s_wsum[w];
  }
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (v & (1u << k)) s_list[off++] = (unsigned short)(threadIdx.x * 4 + k);
  __syncthreads();
  // ---- phase 2: every lane works on a contributing pixel
  const size_t grid_cells = (size_t)a.geo.R * a.geo.W;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int p = base + s_list[i];
    const int r = p / a.geo.W, c = p - r * a.geo.W;
    const float x0 = a.x[((size_t)b * 2) * HW + p];
    const float x1 = a.x[((size_t)b * 2 + 1) * HW + p];
    const float dist = decode_range(x0, a.sigma_mod, a.geo.recip);
    double P[3];
    unproject(dist, a.cos_az[c], a.sin_az[c], a.cos_el[r], a.sin_el[r], P);
    double wx, wy, wz, ww = 1.0;
    if (a.variant == SDPC_VARIANT_POSE) {
      wx = dot4(s_to + 0, P[0], P[1], P[2], 1.0);
      wy = dot4(s_to + 4, P[0], P[1], P[2], 1.0);
      wz = dot4(s_to + 8, P[0], P[1], P[2], 1.0);
      ww = dot4(s_to + 12, P[0], P[1], P[2], 1.0);
    } else {
      wx = P[0] + (double)s_org[src_a * 3 + 0];
      wy = P[1] + (double)s_org[src_a * 3 + 1];
      wz = P[2] + (double)s_org[src_a * 3 + 2];
    }
    const long long inten_fx = inten_to_fixed(x1);
    const unsigned src_id = (unsigned)(src_a * HW + p);
    for (int t = t_lo; t < t_hi; ++t) {
      double qx, qy, qz;
      if (a.variant == SDPC_VARIANT_POSE) {
        const double* m = s_from + (t - t_lo) * 12;
        qx = dot4(m + 0, wx, wy, wz, ww);
        qy = dot4(m + 4, wx, wy, wz, ww);
        qz = dot4(m + 8, wx, wy, wz, ww);
      } else {
        const int ta = t - g * a.A;
        qx = wx - (double)s_org[ta * 3 + 0];
        qy = wy - (double)s_org[ta * 3 + 1];
        qz = wz - (double)s_org[ta * 3 + 2];
      }
      const Candidate cd = reproject_fast(qx, qy, qz, a.sigma_mod, a.geo);
      bool ok = in_grid(cd, a.geo);
      if (a.min_depth_thr >= 0.0f) ok = ok && (cd.nd > (double)a.min_depth_thr);
      if (!ok) continue;
      const size_t cell = (size_t)t * grid_cells + (size_t)cd.row * a.geo.W + cd.col;
      const unsigned long long key = (unsigned long long)__double_as_longlong(cd.nd);
      if constexpr (CAS) {
        zkey_min(a.ws.zkey + cell, key, src_id);
      } else {
        atomicMin(a.ws.zmin + cell, key);
        atomicMin(a.ws.zpack + cell, ((key >> a.key_shift) << a.key_shift) | (unsigned long long)src_id);
      }
      atomicAdd(a.ws.cnt + cell, 1u);
      atomicAdd((unsigned long long*)(a.ws.sum_d + cell), (unsigned long long)depth_to_fixed(cd.nd));
      atomicAdd((unsigned long long*)(a.ws.sum_i + cell), (unsigned long long)inten_fx);
    }
  }
}

// Confirm the packed winners: for every filled cell recompute the exact log-range of the source the packed key
// names; if it is the cell minimum that source is the (smallest-id) nearest candidate, otherwise raise the flag
// that makes the exact winner pass run.
__global__ void __launch_bounds__(256) verify_winner_kernel(ScatterArgs a) {
  const int HW = a.geo.H * a.geo.W;
  const size_t grid_cells = (size_t)a.geo.R * a.geo.W;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)a.tgt_count * grid_cells) return;
  const int t = a.tgt_first + (int)(i / grid_cells);
  const size_t cell = (size_t)a.tgt_first * grid_cells + i;
  if (a.ws.cnt[cell] == 0) return;
  const unsigned id = (unsigned)(a.ws.zpack[cell] & ((1ull << a.key_shift) - 1ull));
  const int g = t / a.A, src_a = id / HW, p = id - src_a * HW, b = g * a.A + src_a;
  const int r = p / a.geo.W, c = p - r * a.geo.W;
  const float dist = decode_range(a.x[((size_t)b * 2) * HW + p], a.sigma_mod, a.geo.recip);
  double P[3], qx, qy, qz;
  unproject(dist, a.cos_az[c], a.sin_az[c], a.cos_el[r], a.sin_el[r], P);
  if (a.variant == SDPC_VARIANT_POSE) {
    const double* tw = a.to_world + (size_t)b * 16;
    const double wx = dot4(tw + 0, P[0], P[1], P[2], 1.0), wy = dot4(tw + 4, P[0], P[1], P[2], 1.0);
    const double wz = dot4(tw + 8, P[0], P[1], P[2], 1.0), ww = dot4(tw + 12, P[0], P[1], P[2], 1.0);
    const double* m = a.from_world + (size_t)t * 16;
    qx = dot4(m + 0, wx, wy, wz, ww);
    qy = dot4(m + 4, wx, wy, wz, ww);
    qz = dot4(m + 8, wx, wy, wz, ww);
  } else {
    const int ta = t - g * a.A;
    qx = (P[0] + (double)a.origins[src_a * 3 + 0]) - (double)a.origins[ta * 3 + 0];
    qy = (P[1] + (double)a.origins[src_a * 3 + 1]) - (double)a.origins[ta * 3 + 1];
    qz = (P[2] + (double)a.origins[src_a * 3 + 2]) - (double)a.origins[ta * 3 + 2];
  }
  const double xy = qx * qx + qy * qy;
  double nd = log2(sqrt(xy + qz * qz) + 1.0);
  nd = sdiv(nd, 6.0, a.geo.recip) * (double)a.sigma_mod;
  if ((unsigned long long)__double_as_longlong(nd) == a.ws.zmin[cell]) a.ws.winner[cell] = id;
  else atomicOr(a.ws.flag, 1u);
}

// ------------------------------------------------------------------------------------------
// resolve
// ------------------------------------------------------------------------------------------
struct ResolveArgs {
  const float* x;
  float* img;                   // [B,2,H,W] newImages (caller's buffer or workspace scratch)
  const int32_t* mask;
  const uint8_t* sky;
  const uint8_t* exist;
  float* new_images;
  int32_t* too_high_out;
  int32_t* dbg_cnt;
  int32_t* dbg_winner;
  double* dbg_min_d;
  StepWorkspace ws;
  GeoConsts geo;
  int A, tgt_first, tgt_count;
  int cas;                      // nearest depth and winner come from ws.zkey (128-bit CAS winner path)
  float sigma_mod, corr_coef;
  double allowance;
};

__global__ void __launch_bounds__(256) resolve_kernel(ResolveArgs a) {
  const int HW = a.geo.H * a.geo.W;
  const int W = a.geo.W, H = a.geo.H, R = a.geo.R;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  const int t = a.tgt_first + blockIdx.y;
  if (p >= HW) return;
  const int r = p / W, c = p - r * W;
  const size_t i0 = ((size_t)t * 2) * HW + p;
  const size_t i1 = i0 + HW;
  const float x0 = a.x[i0];
  const bool neg = x0 < 0.0f;
  // crop rows [R-H, R); negative ranges read the point-mirrored cell (KITTISampling.py:401-403)
  int gr = neg ? (H - 1 - r) : (r + R - H);
  int gc = neg ? ((c - W / 2 + W) % W) : c;
  const size_t cell = (size_t)t * R * W + (size_t)gr * W + gc;
  const unsigned cnt = a.ws.cnt[cell];
  double min_d = 0.0;
  float min_i = 0.0f;
  if (cnt > 0) {
    unsigned w;
    if (a.cas) {
      const ulonglong2 kv = a.ws.zkey[cell];
      min_d = __longlong_as_double((long long)kv.y);
      w = (unsigned)kv.x;
    } else {
      min_d = __longlong_as_double((long long)a.ws.zmin[cell]);
      w = a.ws.winner[cell];
    }
    const int g = t / a.A;
    const int wa = w / HW, wp = w - wa * HW;
    min_i = a.x[((size_t)(g * a.A + wa) * 2 + 1) * HW + wp];
  }
  Fused f = fuse_cell(cnt, a.ws.sum_d[cell], a.ws.sum_i[cell], min_d, min_i, a.sigma_mod, a.allowance, a.geo.recip);
  float nd = (float)(neg ? f.depth * -1.0 : f.depth);
  float ni = f.inten;
  a.img[i0] = nd;
  a.img[i1] = ni;
  a.ws.shared_mask[(size_t)t * HW + p] = (f.filled && (a.exist[p] != 0) && (a.sky[(size_t)t * HW + p] != 0)) ? 1 : 0;
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

// debug dump of the per-cell state (runs before resolve mutates x)
__global__ void dump_cells_kernel(StepWorkspace ws, int32_t* cnt, int32_t* winner, double* min_d, size_t first, size_t n,
                                  int cas) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  size_t k = first + i;
  unsigned cn = ws.cnt[k];
  if (cnt) cnt[k] = (int32_t)cn;
  const ulonglong2 kv = (cas && cn) ? ws.zkey[k] : make_ulonglong2(0ull, 0ull);
  if (winner) winner[k] = cn ? (int32_t)(cas ? (unsigned)kv.x : ws.winner[k]) : -1;
  if (min_d) min_d[k] = cn ? __longlong_as_double((long long)(cas ? kv.y : ws.zmin[k])) : 0.0;
}

}  // namespace sdpc

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
using namespace sdpc;

// The production scatter keeps the exact (depth, source id) minimum with a 128-bit CAS instead of two 64-bit atomicMins
// + verification pass (measured 152.9 -> 131.3 us per 8-view step, same results bit for bit) unless
// SDPC_XVIEW_CAS128=0 (read once) or sdpc_step_params.winner_mode = 1 asks for the packed key.
static bool xview_cas128() {
  static const bool on = [] { const char* v = getenv("SDPC_XVIEW_CAS128"); return !(v && v[0] == '0'); }();
  return on;
}

static int check_params(const sdpc_step_params* p, const sdpc_step_buffers* b) {
  if (!p || !b) return set_error(SDPC_ERR_ARG, "null params/buffers");
  if (p->n_views <= 0 || p->group_size <= 0 || p->n_views % p->group_size != 0)
    return set_error(SDPC_ERR_ARG, "n_views must be a positive multiple of group_size");
  if (p->group_size > kMaxGroup) return set_error(SDPC_ERR_ARG, "group_size > 32 not supported");
  if (p->height <= 0 || p->width <= 0 || (p->height * p->width) % 4 != 0)
    return set_error(SDPC_ERR_ARG, "H*W must be a positive multiple of 4");
  if (p->tgt_first < 0 || p->tgt_count < 0 || p->tgt_first + p->tgt_count > p->n_views)
    return set_error(SDPC_ERR_ARG, "target range outside [0, n_views)");
  if (!b->x || !b->refer || !b->mask) return set_error(SDPC_ERR_ARG, "x/refer/mask must be non-null");
  return SDPC_OK;
}

// the 128-bit CAS winner path serves the production scatter (no candidate-level debug output, H*W a multiple of the chunk)
static bool use_cas128(const sdpc_step_params* p, const sdpc_step_buffers* b) {
  const bool dbg_candidates = b->dbg_row && b->dbg_col && b->dbg_valid;
  const bool fast = !dbg_candidates && (p->height * p->width) % kChunk == 0;
  return fast && (p->winner_mode == 2 || (p->winner_mode == 0 && xview_cas128()));
}

extern "C" int sdpc_step_kernel_launches(const sdpc_step_params* p, const sdpc_step_buffers* b) {
  if (!p || !b) return set_error(SDPC_ERR_ARG, "null params/buffers");
  int n = 1;                                                   // update
  if (p->share) {
    n += 3;                                                    // scatter, resolve, correct
    if (!use_cas128(p, b)) n += 2;                             // verification + exact-winner pass (launched, exits at once)
    if (b->dbg_cnt || b->dbg_winner || b->dbg_min_d) n += 1;   // cell dump
  }
  return n;
}

extern "C" size_t sdpc_step_workspace_bytes(int n_views, int height, int width, int big_rows) {
  return workspace_layout(n_views, height, big_rows, width, nullptr, nullptr);
}

extern "C" int sdpc_langevin_update(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace,
                                    size_t workspace_bytes, void* stream_) {
  if (int e = check_params(p, b)) return e;
  cudaStream_t stream = (cudaStream_t)stream_;
  StepWorkspace ws;
  size_t need = workspace_layout(p->n_views, p->height, p->big_rows, p->width, (char*)workspace, &ws);
  if (!workspace || workspace_bytes < need) return set_error(SDPC_ERR_WORKSPACE, "step workspace too small");
  SDPC_CUDA(cudaMemsetAsync(ws.max_bits, 0, sizeof(unsigned), stream));
  const int HW = p->height * p->width;
  const int tcount = p->tgt_count ? p->tgt_count : p->n_views;
  long long n_vec = (long long)tcount * 2 * HW / 4;
  int blocks = (int)((n_vec + 255) / 256);
  langevin_update_kernel<<<blocks, 256, 0, stream>>>(b->x, b->grad, b->noise, b->refer, b->mask, b->grad_likelihood,
                                                     ws.max_bits, HW, p->tgt_first, n_vec, p->step_size, p->grad_ref,
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
  cudaStream_t stream = (cudaStream_t)stream_;
  StepWorkspace ws;
  size_t need = workspace_layout(p->n_views, p->height, p->big_rows, p->width, (char*)workspace, &ws);
  if (!workspace || workspace_bytes < need) return set_error(SDPC_ERR_WORKSPACE, "step workspace too small");
  const int HW = p->height * p->width;
  const int tcount = p->tgt_count ? p->tgt_count : p->n_views;
  const size_t grid_cells = (size_t)p->big_rows * p->width;
  const size_t first = (size_t)p->tgt_first * grid_cells, n = (size_t)tcount * grid_cells;
  const bool dbg_candidates = b->dbg_row && b->dbg_col && b->dbg_valid;
  const bool fast = !dbg_candidates && HW % kChunk == 0;     // candidate-level debug output: legacy full kernel
  const bool cas = use_cas128(p, b);
  // empty z-buffer: 0xFF.. keys / winners, zero sums and counts (only the target views' grids)
  if (cas) {
    SDPC_CUDA(cudaMemsetAsync(ws.zkey + first, 0xFF, n * 16, stream));
  } else {
    SDPC_CUDA(cudaMemsetAsync(ws.zmin + first, 0xFF, n * 8, stream));
    SDPC_CUDA(cudaMemsetAsync(ws.winner + first, 0xFF, n * 4, stream));
    SDPC_CUDA(cudaMemsetAsync(ws.zpack + first, 0xFF, n * 8, stream));
    SDPC_CUDA(cudaMemsetAsync(ws.flag, 0, sizeof(unsigned), stream));
  }
  SDPC_CUDA(cudaMemsetAsync(ws.sum_d + first, 0, n * 8, stream));
  SDPC_CUDA(cudaMemsetAsync(ws.sum_i + first, 0, n * 8, stream));
  SDPC_CUDA(cudaMemsetAsync(ws.cnt + first, 0, n * 4, stream));

  ScatterArgs sa;
  sa.x = b->x; sa.sky = b->sky; sa.exist = b->exist;
  sa.to_world = b->to_world; sa.from_world = b->from_world; sa.origins = b->origins;
  sa.cos_az = b->cos_az; sa.sin_az = b->sin_az; sa.cos_el = b->cos_el; sa.sin_el = b->sin_el;
  sa.dbg_row = dbg_candidates ? b->dbg_row : nullptr;
  sa.dbg_col = b->dbg_col; sa.dbg_valid = b->dbg_valid;
  sa.ws = ws;
  sa.geo.h_min = p->h_min; sa.geo.dh = p->dh; sa.geo.big_row_min = p->big_row_min; sa.geo.dv = p->dv;
  sa.geo.H = p->height; sa.geo.W = p->width; sa.geo.R = p->big_rows;
  sa.geo.recip = p->scalar_div_recip ? 1 : 0;
  sa.A = p->group_size; sa.variant = p->variant; sa.sky_filter = p->sky_filter;
  sa.tgt_first = p->tgt_first; sa.tgt_count = tcount;
  sa.sigma_mod = p->sigma_mod; sa.min_depth_thr = p->min_depth_thr;
  sa.key_shift = 1;
  while ((1 << sa.key_shift) < p->group_size * HW) ++sa.key_shift;
  if (p->key_shift_override > sa.key_shift && p->key_shift_override < 52) sa.key_shift = p->key_shift_override;
  dim3 grid((HW + 255) / 256, p->group_size, p->n_views / p->group_size);
  if (fast) {
    dim3 fgrid(HW / kChunk, p->group_size, p->n_views / p->group_size);
    if (cas) scatter_fast_kernel<true><<<fgrid, 256, 0, stream>>>(sa);
    else scatter_fast_kernel<false><<<fgrid, 256, 0, stream>>>(sa);
  } else {
    scatter_kernel<0><<<grid, 256, 0, stream>>>(sa);
  }
  SDPC_CUDA(cudaGetLastError());
  if (!cas) {
    verify_winner_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sa);
    SDPC_CUDA(cudaGetLastError());
    scatter_kernel<1><<<grid, 256, 0, stream>>>(sa);            // exits immediately unless a winner was unconfirmed
    SDPC_CUDA(cudaGetLastError());
  }
  if (b->dbg_cnt || b->dbg_winner || b->dbg_min_d) {
    dump_cells_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(ws, b->dbg_cnt, b->dbg_winner, b->dbg_min_d, first, n,
                                                                      cas ? 1 : 0);
    SDPC_CUDA(cudaGetLastError());
  }
  ResolveArgs ra;
  ra.x = b->x; ra.mask = b->mask; ra.sky = b->sky; ra.exist = b->exist;
  ra.img = b->new_images ? b->new_images : ws.shared_img;
  ra.new_images = b->new_images; ra.too_high_out = b->too_high;
  ra.dbg_cnt = b->dbg_cnt; ra.dbg_winner = b->dbg_winner; ra.dbg_min_d = b->dbg_min_d;
  ra.ws = ws; ra.geo = sa.geo; ra.A = p->group_size; ra.tgt_first = p->tgt_first; ra.tgt_count = tcount;
  ra.cas = cas ? 1 : 0;
  ra.sigma_mod = p->sigma_mod; ra.corr_coef = p->corr_coef; ra.allowance = p->allowance;
  dim3 rgrid((HW + 255) / 256, tcount);
  resolve_kernel<<<rgrid, 256, 0, stream>>>(ra);
  SDPC_CUDA(cudaGetLastError());
  long long n_vec = (long long)tcount * 2 * HW / 4;
  correct_kernel<<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(b->x, ra.img, ws.shared_mask, b->mask, ws.max_bits,
                                                                     b->too_high, HW, p->tgt_first, n_vec,
                                                                     p->sigma_mod, p->corr_coef, sa.geo.recip);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_langevin_reproject_step(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace,
                                            size_t workspace_bytes, void* stream) {
  if (int e = sdpc_langevin_update(p, b, workspace, workspace_bytes, stream)) return e;
  if (p->share) return sdpc_crossview_share(p, b, workspace, workspace_bytes, stream);
  return SDPC_OK;
}

extern "C" int sdpc_langevin_reproject_step_host(const sdpc_step_params* p, const sdpc_step_buffers* b, float* x_host,
                                                 const float* grad_host, const float* noise_host,
                                                 float* new_images_host, void* workspace, size_t workspace_bytes,
                                                 void* stream_) {
  if (int e = check_params(p, b)) return e;
  if (!x_host) return set_error(SDPC_ERR_ARG, "x_host is null");
  cudaStream_t stream = (cudaStream_t)stream_;
  const size_t bytes = (size_t)p->n_views * 2 * p->height * p->width * sizeof(float);
  SDPC_CUDA(cudaMemcpyAsync(b->x, x_host, bytes, cudaMemcpyHostToDevice, stream));
  if (grad_host) {
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
