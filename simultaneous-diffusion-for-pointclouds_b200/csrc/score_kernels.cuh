// CUDA-core kernels of the score network: everything that is not a 128/256-channel convolution.
// (begin/end convolutions, InstanceNorm++ statistics, operand materialisation with halo,
// max / mean pooling, bilinear x2, layout conversion, fp32 SIMT convolution for the strict arm.)
#pragma once
#include <type_traits>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.h"
#include "score_types.cuh"

namespace sdpc {

// ------------------------------------------------------------------------------------------
// begin_conv: 4->ngf 3x3 zero-padded (ncsnv2.py:486-498), input built on the fly:
// channels 0,1 = 2x-1, channel 2 = linspace(0,1,W)[w], channel 3 = linspace(0,1,H)[h].
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float linspace01(int i, int steps) {
  // torch.linspace float32 kernel: symmetric evaluation around the midpoint
  const float step = 1.0f / (float)(steps - 1);
  return (i < steps / 2) ? step * (float)i : 1.0f - step * (float)(steps - 1 - i);
}

template <int NGF>
__global__ void __launch_bounds__(NGF)
begin_conv_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                  float* __restrict__ out, float2* __restrict__ stat_parts, int N, int H, int W) {
  pdl_sync();
  // one thread per output channel (its 36 weights live in registers), 64 pixels of one row per block;
  // the 36-value input patches are staged k-major in shared memory and read as broadcast float4s.
  // stat_parts (or null): [N][H * W/64][NGF] (sum, sum of squares) of the block's 64 outputs per channel, the
  // partial InstanceNorm++ statistics of the first normalisation (same slot layout as the conv epilogues).
  constexpr int K = 36;
  __shared__ __align__(16) float sin_[K][64];
  const int tiles_w = W / 64;
  const int blk = blockIdx.x;
  const int w0 = (blk % tiles_w) * 64;
  const int h = (blk / tiles_w) % H;
  const int n = blk / (tiles_w * H);
  const int c = threadIdx.x;
  float wr[K];
#pragma unroll
  for (int k = 0; k < K; ++k) wr[k] = wgt[c * K + k];          // wgt is [co][ci][kh][kw] = [co][k]
  const float bc = bias[c];
  for (int i = threadIdx.x; i < 64 * K; i += NGF) {
    const int k = i / 64, p = i % 64;
    const int ci = k / 9, kh = (k % 9) / 3, kw = k % 3;
    const int hh = h + kh - 1, ww = w0 + p + kw - 1;
    float v = 0.0f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      if (ci < 2) v = 2.0f * x[(((size_t)n * 2 + ci) * H + hh) * W + ww] - 1.0f;
      else if (ci == 2) v = linspace01(ww, W);
      else v = linspace01(hh, H);
    }
    sin_[k][p] = v;
  }
  __syncthreads();
  float* orow = out + (((size_t)n * H + h) * W + w0) * NGF + c;
  float ssum = 0.0f, ssq = 0.0f;
#pragma unroll 1
  for (int p0 = 0; p0 < 64; p0 += 8) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&sin_[k][p0]);
      const float4 b = *reinterpret_cast<const float4*>(&sin_[k][p0 + 4]);
      acc[0] = fmaf(a.x, wr[k], acc[0]); acc[1] = fmaf(a.y, wr[k], acc[1]);
      acc[2] = fmaf(a.z, wr[k], acc[2]); acc[3] = fmaf(a.w, wr[k], acc[3]);
      acc[4] = fmaf(b.x, wr[k], acc[4]); acc[5] = fmaf(b.y, wr[k], acc[5]);
      acc[6] = fmaf(b.z, wr[k], acc[6]); acc[7] = fmaf(b.w, wr[k], acc[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float v = acc[i] + bc;
      orow[(size_t)(p0 + i) * NGF] = v;
      ssum += v;
      ssq = fmaf(v, v, ssq);
    }
  }
  if (stat_parts) stat_parts[(size_t)blk * NGF + c] = make_float2(ssum, ssq);
}

// ------------------------------------------------------------------------------------------
// InstanceNorm++ (normalization.py:163-176)
// ------------------------------------------------------------------------------------------
// per-(n,c) sum and sum of squares over H*W; fp32 partials per thread, fp64 across threads.
__global__ void __launch_bounds__(256)
stats_kernel(const float* __restrict__ in, double* __restrict__ stats, int HW, int C, int pix_per_block) {
  pdl_sync();
  extern __shared__ double sred[];                  // [groups][C][2]
  const int lanes_c = C / 4;
  const int groups = blockDim.x / lanes_c;
  const int cq = threadIdx.x % lanes_c, grp = threadIdx.x / lanes_c;
  const int n = blockIdx.y;
  const size_t p0 = (size_t)blockIdx.x * pix_per_block;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
  for (int p = grp; p < pix_per_block; p += groups) {
    const float4 v = *reinterpret_cast<const float4*>(in + ((size_t)n * HW + p0 + p) * C + cq * 4);
    s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    q.x = fmaf(v.x, v.x, q.x); q.y = fmaf(v.y, v.y, q.y); q.z = fmaf(v.z, v.z, q.z); q.w = fmaf(v.w, v.w, q.w);
  }
  double* mine = sred + ((size_t)grp * C + cq * 4) * 2;
  mine[0] = s.x; mine[1] = q.x; mine[2] = s.y; mine[3] = q.y; mine[4] = s.z; mine[5] = q.z; mine[6] = s.w; mine[7] = q.w;
  __syncthreads();
  for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
    double acc = 0.0;
    for (int g2 = 0; g2 < groups; ++g2) acc += sred[(size_t)g2 * C * 2 + i];
    atomicAdd(stats + (size_t)n * C * 2 + i, acc);
  }
}

// coef[n][c] = {mean, a, b}: out = a*(x-mean) + b with a = gamma*rstd, b = gamma*alpha*mean_n + beta
// (InstanceNorm2dPlus, normalization.py:163-176).  Called by the first C threads of a block (C <= 256); `S`, `Q` are
// the channel's sum and sum of squares over the image.
// block-wide sum of one double per thread (blockDim.x a multiple of 32, <= 256), fixed reduction tree
__device__ __forceinline__ double block_sum_256(double v, double* warp_part /*[8]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                    // warp_part may still be read from a previous call
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  const int nw = blockDim.x >> 5;
  for (int i = 0; i < nw; ++i) t += warp_part[i];
  return t;
}

__device__ __forceinline__ void norm_coefficients(double S, double Q, const float* __restrict__ alpha,
                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                  float* __restrict__ coef, int HW, int C, int n, int c, bool active) {
  __shared__ double wp[8];
  const double mean = S / HW;
  double var = Q / HW - mean * mean;
  if (var < 0.0) var = 0.0;
  const double s_m = block_sum_256(active ? mean : 0.0, wp) / C;
  const double dm = mean - s_m;
  const double s_v = block_sum_256(active ? dm * dm : 0.0, wp) / (C - 1);
  if (!active) return;
  const float rstd = 1.0f / sqrtf((float)var + 1e-5f);
  const float mean_n = (float)dm / sqrtf((float)s_v + 1e-5f);
  float* o = coef + ((size_t)n * C + c) * 3;
  o[0] = (float)mean;
  o[1] = gamma[c] * rstd;
  o[2] = gamma[c] * (mean_n * alpha[c]) + beta[c];
}

__global__ void norm_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ alpha,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ coef, int HW, int C) {
  pdl_sync();
  const int n = blockIdx.x, c = threadIdx.x;
  norm_coefficients(stats[((size_t)n * C + c) * 2], stats[((size_t)n * C + c) * 2 + 1], alpha, gamma, beta, coef, HW, C, n, c,
                    true);
}

// Fused reducer + finalizer for the statistics the convolution epilogues leave as per-tile partial sums
// (parts: [N][parts_per_view][C][2] floats): grid (C/8, N), block 256 = 32 part-groups x 8 channels.  Every block
// sums its 8 channels into stats [N][C][2] (doubles); the block that finishes an image last (per-image ticket
// counter, reset for the next use) turns the image's C sums into the normalisation coefficients.
__global__ void __launch_bounds__(256)
stats_reduce_finalize_kernel(const float* __restrict__ parts, double* __restrict__ stats, int parts_per_view, int C,
                             const float* __restrict__ alpha, const float* __restrict__ gamma,
                             const float* __restrict__ beta, float* __restrict__ coef, int HW,
                             unsigned int* __restrict__ tickets) {
  pdl_sync();
  __shared__ double red[8][8][2];
  __shared__ bool s_last;
  const int cl = threadIdx.x & 7, grp = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl, n = blockIdx.y;
  const float2* p = reinterpret_cast<const float2*>(parts) + (size_t)n * parts_per_view * C + c;
  double S = 0.0, Q = 0.0;
#pragma unroll 8
  for (int j = grp; j < parts_per_view; j += 32) {
    const float2 v = p[(size_t)j * C];
    S += (double)v.x;
    Q += (double)v.y;
  }
  // lanes of a warp hold 4 part-groups x 8 channels: fold the groups with shuffles, then the 8 warps through shared memory
  S += __shfl_xor_sync(0xffffffffu, S, 8);  Q += __shfl_xor_sync(0xffffffffu, Q, 8);
  S += __shfl_xor_sync(0xffffffffu, S, 16); Q += __shfl_xor_sync(0xffffffffu, Q, 16);
  if ((threadIdx.x & 31) < 8) { red[threadIdx.x >> 5][cl][0] = S; red[threadIdx.x >> 5][cl][1] = Q; }
  __syncthreads();
  if (threadIdx.x < 8) {
    S = red[0][cl][0]; Q = red[0][cl][1];
#pragma unroll
    for (int k = 1; k < 8; ++k) { S += red[k][cl][0]; Q += red[k][cl][1]; }
    stats[((size_t)n * C + c) * 2] = S;
    stats[((size_t)n * C + c) * 2 + 1] = Q;
    __threadfence();
  }
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(tickets + n, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const int t = threadIdx.x;
  const bool active = t < C;
  double fS = 0.0, fQ = 0.0;
  if (active) {
    fS = __ldcg(stats + ((size_t)n * C + t) * 2);
    fQ = __ldcg(stats + ((size_t)n * C + t) * 2 + 1);
  }
  norm_coefficients(fS, fQ, alpha, gamma, beta, coef, HW, C, n, t, active);
  if (t == 0) tickets[n] = 0;
}

// ------------------------------------------------------------------------------------------
// operand materialisation: raw fp32 NHWC -> T NHWC with halo
// ------------------------------------------------------------------------------------------
enum { OP_COPY = 0, OP_ELU = 1, OP_NORM_ELU = 2 };
enum { HALO_CIRC = 0, HALO_ZERO = 1 };

template <typename T>
__device__ __forceinline__ void store_op8(T* dst, const float* v, bool tf32, size_t lo_off = 0) {
  store_op4<T>(dst, v, tf32, lo_off);
  store_op4<T>(dst + 4, v + 4, tf32, lo_off);
}
// bf16: one 16-byte store per plane instead of two 8-byte ones
template <>
__device__ __forceinline__ void store_op8<__nv_bfloat16>(__nv_bfloat16* dst, const float* v, bool alt, size_t lo_off) {
  if (alt) {                                                  // fp16 arm
    *reinterpret_cast<uint4*>(dst) = make_uint4(pack2_h16(v[0], v[1], true), pack2_h16(v[2], v[3], true),
                                                pack2_h16(v[4], v[5], true), pack2_h16(v[6], v[7], true));
    return;
  }
  __nv_bfloat162 h[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) h[k] = __floats2bfloat162_rn(v[2 * k], v[2 * k + 1]);
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
  if (lo_off) {
    __nv_bfloat162 l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h[k]);
      l[k] = __floats2bfloat162_rn(v[2 * k] - f.x, v[2 * k + 1] - f.y);
    }
    *reinterpret_cast<uint4*>(dst + lo_off) = *reinterpret_cast<const uint4*>(l);
  }
}

// One thread: 8 channels x PIX consecutive pixels of a row (all loads issued before use: 128 bytes in flight per
// thread, i.e. 4 pixels of fp32 input or 8 pixels of bf16 input).  With HALO_ZERO the halo is not written here: the
// buffer's border is cleared by zero_halo_kernel.
#ifndef SDPC_OP_PIX_F32            // tuning knobs of tools/build_variant.py -D ...: pixels per thread, resident blocks per SM
#define SDPC_OP_PIX_F32 4
#endif
#ifndef SDPC_OP_PIX_H16
#define SDPC_OP_PIX_H16 8
#endif
#ifndef SDPC_OP_MINB
#define SDPC_OP_MINB 2
#endif
template <typename TIn> __host__ __device__ constexpr int op_pix() { return sizeof(TIn) == 2 ? SDPC_OP_PIX_H16 : SDPC_OP_PIX_F32; }
template <typename T, int MINB, typename TIn = float>
__global__ void __launch_bounds__(256, MINB)
to_operand_kernel(const TIn* __restrict__ in, const float* __restrict__ coef, T* __restrict__ out, int N, int H,
                  int W, int C, int P, int mode, int halo, int tf32, size_t lo_off) {
  pdl_sync();
  constexpr int PIX = op_pix<TIn>();
  const int C8 = C / 8, WS = W / PIX;
  const size_t total = (size_t)N * H * WS * C8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  size_t t = i / C8;
  const int w0 = (int)(t % WS) * PIX; t /= WS;
  const int h = (int)(t % H);
  const int n = (int)(t / H);
  const int Hp = H + 2 * P, Wp = W + 2 * P;
  const TIn* src = in + (((size_t)n * H + h) * W + w0) * C + c8 * 8;
  uint4 q[PIX], q2[PIX];                                       // raw input words: q (+ q2 for the second half of fp32 input)
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    q[j] = *reinterpret_cast<const uint4*>(src + (size_t)j * C);
    if constexpr (sizeof(TIn) == 4) q2[j] = *reinterpret_cast<const uint4*>(src + (size_t)j * C + 4);
  }
  float mu[8], ga[8], be[8];
  if (mode == OP_NORM_ELU) {
    // 8 channels x {mean, a, b} = 24 contiguous floats: six 16-byte loads
    const float4* cf4 = reinterpret_cast<const float4*>(coef + ((size_t)n * C + c8 * 8) * 3);
    float cf[24];
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const float4 t4 = cf4[k];
      cf[4 * k] = t4.x; cf[4 * k + 1] = t4.y; cf[4 * k + 2] = t4.z; cf[4 * k + 3] = t4.w;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) { mu[k] = cf[k * 3]; ga[k] = cf[k * 3 + 1]; be[k] = cf[k * 3 + 2]; }
  }
#pragma unroll
  for (int j = 0; j < PIX; ++j) {
    float v[8];
    if constexpr (sizeof(TIn) == 2) {                          // bf16 input (conv1 outputs of the plain bf16 arm)
      const float2 f0 = unpack2_h16(q[j].x, tf32 != 0), f1 = unpack2_h16(q[j].y, tf32 != 0);
      const float2 f2 = unpack2_h16(q[j].z, tf32 != 0), f3 = unpack2_h16(q[j].w, tf32 != 0);
      v[0] = f0.x; v[1] = f0.y; v[2] = f1.x; v[3] = f1.y; v[4] = f2.x; v[5] = f2.y; v[6] = f3.x; v[7] = f3.y;
    } else {
      v[0] = __uint_as_float(q[j].x); v[1] = __uint_as_float(q[j].y); v[2] = __uint_as_float(q[j].z); v[3] = __uint_as_float(q[j].w);
      v[4] = __uint_as_float(q2[j].x); v[5] = __uint_as_float(q2[j].y); v[6] = __uint_as_float(q2[j].z); v[7] = __uint_as_float(q2[j].w);
    }
    if (mode == OP_NORM_ELU) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = elu_sel<T>(ga[k] * (v[k] - mu[k]) + be[k], tf32 != 0);
    } else if (mode == OP_ELU) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = elu_sel<T>(v[k], tf32 != 0);
    }
    const int w = w0 + j;
    if (halo == HALO_ZERO) {
      store_op8<T>(out + (((size_t)n * Hp + h + P) * Wp + w + P) * C + c8 * 8, v, tf32 != 0, lo_off);
    } else {
      const HaloPos d = halo_pos(h, w, H, W, P);
      for_each_halo_pos(d, [&](int hp, int wp) {
        store_op8<T>(out + (((size_t)n * Hp + hp) * Wp + wp) * C + c8 * 8, v, tf32 != 0, lo_off);
      });
    }
  }
}

// zero the P-wide border of a padded NHWC tensor (zero-padded convolutions)
template <typename T>
__global__ void __launch_bounds__(256)
zero_halo_kernel(T* __restrict__ out, int N, int H, int W, int C, int P) {
  pdl_sync();
  const int Hp = H + 2 * P, Wp = W + 2 * P, C8 = C / 8;
  const int border = Hp * Wp - H * W;                 // border pixels per image
  const size_t total = (size_t)N * border * C8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  size_t t = i / C8;
  int bidx = (int)(t % border);
  const int n = (int)(t / border);
  int hp, wp;
  const int top = P * Wp;
  if (bidx < top) { hp = bidx / Wp; wp = bidx % Wp; }
  else if (bidx < 2 * top) { bidx -= top; hp = H + P + bidx / Wp; wp = bidx % Wp; }
  else { bidx -= 2 * top; hp = P + bidx / (2 * P); const int q = bidx % (2 * P); wp = q < P ? q : W + q; }
  float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  store_op8<T>(out + (((size_t)n * Hp + hp) * Wp + wp) * C + c8 * 8, z, false);
}

// CRP stage (layers.py:76-83): out_op = [ELU](maxpool5(in)) with circular halo for the conv that
// follows; MaxPool2d(5,1,2) itself pads with -inf, i.e. the window is clipped at the image border.
// ELU is monotone, so ELU(maxpool(x)) == maxpool(ELU(x)).  Optionally also emits x0 = ELU(in).
constexpr int kPoolTH = 8, kPoolTW = 32, kPoolCB = 32;     // output tile: 8 rows x 32 columns x 32 channels per block
constexpr int kPoolThreads = (kPoolTW + 4) * (kPoolCB / 4);  // 288: one thread per (input column, float4 of channels)
constexpr int kPoolSmemBytes = kPoolTH * (kPoolTW + 4) * (kPoolCB / 4) * 16;   // 36 KB

template <typename T>
__global__ void __launch_bounds__(kPoolThreads)
maxpool5_kernel(const float* __restrict__ in, float* __restrict__ x0_out, T* __restrict__ out, int N, int H, int W,
                int C, int P, int elu_in, int tf32, size_t lo_off) {
  pdl_sync();
  // Separable 5x5 max (window clipped at the image border = MaxPool2d's -inf padding).
  //   phase A: a thread owns one input column (of the 32+4) x 4 channels, loads its 8+4 rows straight from global
  //            memory (independent 128-bit loads, 8 lanes = one pixel's 128 contiguous bytes) and keeps a running
  //            vertical 5-max in registers; only the 8 results go to shared memory.  The same thread emits
  //            x0 = ELU(in) for its interior rows.
  //   phase B: a thread owns 8 consecutive outputs of one row x 4 channels and slides the horizontal 5-max over the
  //            12 shared-memory columns it needs, then stores the operand with its circular-halo duplicates.
  __shared__ float4 tv[kPoolTH][kPoolTW + 4][kPoolCB / 4];
  const int CBn = C / kPoolCB, TWn = W / kPoolTW, THn = H / kPoolTH;
  int b = blockIdx.x;
  const int cb = b % CBn; b /= CBn;
  const int tw = b % TWn; b /= TWn;
  const int th = b % THn;
  const int n = b / THn;
  const int h0 = th * kPoolTH, w0 = tw * kPoolTW, c0 = cb * kPoolCB;
  const bool red = tf32 != 0;
  const float4 ninf = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
  {
    const int c4 = threadIdx.x & 7, cc = threadIdx.x >> 3;          // cc in [0, 36)
    const int ww = w0 - 2 + cc;
    const bool col_ok = ww >= 0 && ww < W;
    float4 v[kPoolTH + 4];
#pragma unroll
    for (int r = 0; r < kPoolTH + 4; ++r) {
      const int hh = h0 - 2 + r;
      v[r] = ninf;
      if (col_ok && hh >= 0 && hh < H)
        v[r] = *reinterpret_cast<const float4*>(in + (((size_t)n * H + hh) * W + ww) * C + c0 + c4 * 4);
    }
#pragma unroll
    for (int r = 0; r < kPoolTH; ++r) {
      float4 m = v[r];
#pragma unroll
      for (int k = 1; k < 5; ++k) {
        m.x = fmaxf(m.x, v[r + k].x); m.y = fmaxf(m.y, v[r + k].y); m.z = fmaxf(m.z, v[r + k].z); m.w = fmaxf(m.w, v[r + k].w);
      }
      tv[r][cc][c4] = m;
    }
    if (x0_out && cc >= 2 && cc < kPoolTW + 2) {
#pragma unroll
      for (int r = 0; r < kPoolTH; ++r) {
        const float4 a = v[r + 2];
        *reinterpret_cast<float4*>(x0_out + (((size_t)n * H + h0 + r) * W + ww) * C + c0 + c4 * 4) =
            make_float4(elu_sel<T>(a.x, red), elu_sel<T>(a.y, red), elu_sel<T>(a.z, red), elu_sel<T>(a.w, red));
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    const int c4 = threadIdx.x & 7, seg = (threadIdx.x >> 3) & 3, r = threadIdx.x >> 5;   // 8 outputs: columns seg*8 ..
    const int Hp = H + 2 * P, Wp = W + 2 * P;
    const int h = h0 + r;
    float4 cfl[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) cfl[j] = tv[r][seg * 8 + j][c4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 m = cfl[j];
#pragma unroll
      for (int k = 1; k < 5; ++k) {
        m.x = fmaxf(m.x, cfl[j + k].x); m.y = fmaxf(m.y, cfl[j + k].y); m.z = fmaxf(m.z, cfl[j + k].z); m.w = fmaxf(m.w, cfl[j + k].w);
      }
      float o[4] = {m.x, m.y, m.z, m.w};
      if (elu_in) { o[0] = elu_sel<T>(o[0], red); o[1] = elu_sel<T>(o[1], red); o[2] = elu_sel<T>(o[2], red); o[3] = elu_sel<T>(o[3], red); }
      const int w = w0 + seg * 8 + j;
      const HaloPos d = halo_pos(h, w, H, W, P);
      for_each_halo_pos(d, [&](int hp, int wp) {
        store_op4<T>(out + (((size_t)n * Hp + hp) * Wp + wp) * C + c0 + c4 * 4, o, red, lo_off);
      });
    }
  }
}

// bf16 arm of the CRP max-pool: rounding to bf16 is monotone, so bf16(max(x)) == max(bf16(x)) and the whole 5x5 max
// runs on packed bf16 pairs (half the instructions and half the shared-memory bytes of the fp32 kernel above).
// A thread owns 8 channels (one 16-byte word); the block tile is 8 rows x 32 columns x 64 channels.  The input is
// either the fp32 trunk tensor (first stage: ELU applied at load, x0 = ELU(in) emitted in fp32 for the residual
// adds) or the bf16 copy the previous CRP convolution left in its out_acc (second stage).
constexpr int kPoolHCB = 64;
constexpr int kPoolHThreads = (kPoolTW + 4) * (kPoolHCB / 8);   // 288

__device__ __forceinline__ uint32_t hmax2_half(uint32_t a, uint32_t b) {
  const __half2 m = __hmax2(*reinterpret_cast<const __half2*>(&a), *reinterpret_cast<const __half2*>(&b));
  return *reinterpret_cast<const uint32_t*>(&m);
}
__device__ __forceinline__ uint4 hmax8(const uint4& a, const uint4& b, bool alt) {
  uint4 r;
  if (alt) {                                                  // fp16 arm: the words are half pairs
    r.x = hmax2_half(a.x, b.x); r.y = hmax2_half(a.y, b.y); r.z = hmax2_half(a.z, b.z); r.w = hmax2_half(a.w, b.w);
    return r;
  }
  const __nv_bfloat162 x = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.x), *reinterpret_cast<const __nv_bfloat162*>(&b.x));
  const __nv_bfloat162 y = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.y), *reinterpret_cast<const __nv_bfloat162*>(&b.y));
  const __nv_bfloat162 z = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.z), *reinterpret_cast<const __nv_bfloat162*>(&b.z));
  const __nv_bfloat162 w = __hmax2(*reinterpret_cast<const __nv_bfloat162*>(&a.w), *reinterpret_cast<const __nv_bfloat162*>(&b.w));
  r.x = *reinterpret_cast<const uint32_t*>(&x); r.y = *reinterpret_cast<const uint32_t*>(&y);
  r.z = *reinterpret_cast<const uint32_t*>(&z); r.w = *reinterpret_cast<const uint32_t*>(&w);
  return r;
}

template <typename TIn>
__global__ void __launch_bounds__(kPoolHThreads)
maxpool5_h2_kernel(const TIn* __restrict__ in, float* __restrict__ x0_out, __nv_bfloat16* __restrict__ out, int N, int H,
                   int W, int C, int P, int elu_in, int alt) {
  pdl_sync();
  __shared__ uint4 tv[kPoolTH][kPoolTW + 4][kPoolHCB / 8];       // 36 KB
  const int CBn = C / kPoolHCB, TWn = W / kPoolTW, THn = H / kPoolTH;
  int b = blockIdx.x;
  const int cb = b % CBn; b /= CBn;
  const int tw = b % TWn; b /= TWn;
  const int th = b % THn;
  const int n = b / THn;
  const int h0 = th * kPoolTH, w0 = tw * kPoolTW, c0 = cb * kPoolHCB;
  const uint32_t ni2 = alt ? 0xFC00FC00u : 0xFF80FF80u;                                   // -inf pairs (half / bf16)
  const uint4 ninf = make_uint4(ni2, ni2, ni2, ni2);
  const bool h16 = alt != 0;
  {
    const int c8 = threadIdx.x & 7, cc = threadIdx.x >> 3;        // cc in [0, 36): input column
    const int ww = w0 - 2 + cc;
    const bool col_ok = ww >= 0 && ww < W;
    const bool interior_col = cc >= 2 && cc < kPoolTW + 2;
    uint4 v[kPoolTH + 4];
#pragma unroll
    for (int r = 0; r < kPoolTH + 4; ++r) {
      const int hh = h0 - 2 + r;
      v[r] = ninf;
      if (col_ok && hh >= 0 && hh < H) {
        const size_t off = (((size_t)n * H + hh) * W + ww) * C + c0 + c8 * 8;
        if constexpr (sizeof(TIn) == 2) {
          v[r] = *reinterpret_cast<const uint4*>(in + off);
        } else {
          float4 a = *reinterpret_cast<const float4*>(in + off), c = *reinterpret_cast<const float4*>(in + off + 4);
          if (elu_in || x0_out) {
            const float4 ea = make_float4(elu_fast(a.x), elu_fast(a.y), elu_fast(a.z), elu_fast(a.w));
            const float4 ec = make_float4(elu_fast(c.x), elu_fast(c.y), elu_fast(c.z), elu_fast(c.w));
            if (x0_out && interior_col && r >= 2 && r < kPoolTH + 2) {
              *reinterpret_cast<float4*>(x0_out + off) = ea;
              *reinterpret_cast<float4*>(x0_out + off + 4) = ec;
            }
            if (elu_in) { a = ea; c = ec; }
          }
          v[r] = make_uint4(pack2_h16(a.x, a.y, h16), pack2_h16(a.z, a.w, h16), pack2_h16(c.x, c.y, h16), pack2_h16(c.z, c.w, h16));
        }
      }
    }
#pragma unroll
    for (int r = 0; r < kPoolTH; ++r) {
      uint4 m = v[r];
#pragma unroll
      for (int k = 1; k < 5; ++k) m = hmax8(m, v[r + k], h16);
      tv[r][cc][c8] = m;
    }
  }
  __syncthreads();
  if (threadIdx.x < 256) {
    const int c8 = threadIdx.x & 7, seg = (threadIdx.x >> 3) & 3, r = threadIdx.x >> 5;   // 8 outputs: columns seg*8 ..
    const int Hp = H + 2 * P, Wp = W + 2 * P;
    const int h = h0 + r;
    uint4 cfl[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) cfl[j] = tv[r][seg * 8 + j][c8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint4 m = cfl[j];
#pragma unroll
      for (int k = 1; k < 5; ++k) m = hmax8(m, cfl[j + k], h16);
      const int w = w0 + seg * 8 + j;
      const HaloPos d = halo_pos(h, w, H, W, P);
      for_each_halo_pos(d, [&](int hp, int wp) {
        *reinterpret_cast<uint4*>(out + (((size_t)n * Hp + hp) * Wp + wp) * C + c0 + c8 * 8) = m;
      });
    }
  }
}

// mean of the four stride-2 phases (layers.py:310-312), in the reference's summation order.
__device__ __forceinline__ float4 pool4(const float* in, size_t i00, size_t i10, size_t i01, size_t i11) {
  const float4 a = *reinterpret_cast<const float4*>(in + i00), b = *reinterpret_cast<const float4*>(in + i10);
  const float4 c = *reinterpret_cast<const float4*>(in + i01), d = *reinterpret_cast<const float4*>(in + i11);
  return make_float4((((a.x + b.x) + c.x) + d.x) / 4.0f, (((a.y + b.y) + c.y) + d.y) / 4.0f,
                     (((a.z + b.z) + c.z) + d.z) / 4.0f, (((a.w + b.w) + c.w) + d.w) / 4.0f);
}

// out (T, no halo) = meanpool2(in); out_add (fp32) = meanpool2(in) + add
template <typename T>
__global__ void __launch_bounds__(256)
meanpool_kernel(const float* __restrict__ in, const float* __restrict__ add, T* __restrict__ out_op,
                float* __restrict__ out_raw, int N, int H, int W, int C, int tf32, size_t lo_off) {
  pdl_sync();
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const size_t total = (size_t)N * Ho * Wo * C4;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  size_t t = i / C4;
  const int wo = (int)(t % Wo); t /= Wo;
  const int ho = (int)(t % Ho);
  const int n = (int)(t / Ho);
  const size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c4 * 4;
  float4 v = pool4(in, base, base + (size_t)W * C, base + C, base + (size_t)W * C + C);
  if (add) {
    const float4 a = *reinterpret_cast<const float4*>(add + i * 4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
  }
  if (out_raw) *reinterpret_cast<float4*>(out_raw + i * 4) = v;
  if (out_op) {
    float vv[4] = {v.x, v.y, v.z, v.w};
    store_op4<T>(out_op + i * 4, vv, tf32 != 0, lo_off);
  }
}

// out_raw = meanpool2(in) + add, plus partial InstanceNorm++ statistics of out_raw: a block covers kMpIter * 256 / (C/4)
// consecutive output pixels of one image and leaves (sum, sum of squares) per channel in its slot
// stat_parts[block][C] (the [N][parts_per_view][C][2] layout of the convolution epilogues).  Needs 256 % (C/4) == 0.
constexpr int kMpIter = 16;
__global__ void __launch_bounds__(256)
meanpool_add_stats_kernel(const float* __restrict__ in, const float* __restrict__ add, float* __restrict__ out_raw,
                          float2* __restrict__ stat_parts, int N, int H, int W, int C) {
  pdl_sync();
  __shared__ float4 ssum[256], ssq[256];
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const int c4 = threadIdx.x % C4;
  float4 S = make_float4(0.f, 0.f, 0.f, 0.f), Q = S;
#pragma unroll 4
  for (int it = 0; it < kMpIter; ++it) {
    const size_t i = ((size_t)blockIdx.x * kMpIter + it) * 256 + threadIdx.x;
    size_t t = i / C4;
    const int wo = (int)(t % Wo); t /= Wo;
    const int ho = (int)(t % Ho);
    const int n = (int)(t / Ho);
    const size_t base = (((size_t)n * H + 2 * ho) * W + 2 * wo) * C + c4 * 4;
    float4 v = pool4(in, base, base + (size_t)W * C, base + C, base + (size_t)W * C + C);
    const float4 a = *reinterpret_cast<const float4*>(add + i * 4);
    v.x += a.x; v.y += a.y; v.z += a.z; v.w += a.w;
    *reinterpret_cast<float4*>(out_raw + i * 4) = v;
    S.x += v.x; S.y += v.y; S.z += v.z; S.w += v.w;
    Q.x = fmaf(v.x, v.x, Q.x); Q.y = fmaf(v.y, v.y, Q.y); Q.z = fmaf(v.z, v.z, Q.z); Q.w = fmaf(v.w, v.w, Q.w);
  }
  ssum[threadIdx.x] = S;
  ssq[threadIdx.x] = Q;
  __syncthreads();
  if (threadIdx.x < C4) {
    for (int k = threadIdx.x + C4; k < 256; k += C4) {
      const float4 s2 = ssum[k], q2 = ssq[k];
      S.x += s2.x; S.y += s2.y; S.z += s2.z; S.w += s2.w;
      Q.x += q2.x; Q.y += q2.y; Q.z += q2.z; Q.w += q2.w;
    }
    float2* sp = stat_parts + (size_t)blockIdx.x * C + c4 * 4;
    sp[0] = make_float2(S.x, Q.x); sp[1] = make_float2(S.y, Q.y); sp[2] = make_float2(S.z, Q.z); sp[3] = make_float2(S.w, Q.w);
  }
}

// out = a + bilinear_x2(b), align_corners=True (layers.py:182; F.interpolate)
__global__ void __launch_bounds__(256)
upsample_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int N, int H,
                    int W, int C, int h, int w) {
  pdl_sync();
  const int C4 = C / 4;
  const size_t total = (size_t)N * H * W * C4;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  size_t t = i / C4;
  const int x = (int)(t % W); t /= W;
  const int y = (int)(t % H);
  const int n = (int)(t / H);
  const float sh = (H > 1) ? (float)(h - 1) / (float)(H - 1) : 0.0f;
  const float sw = (W > 1) ? (float)(w - 1) / (float)(W - 1) : 0.0f;
  const float fy = sh * (float)y, fx = sw * (float)x;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < h - 1 ? 1 : 0), x1 = x0 + (x0 < w - 1 ? 1 : 0);
  const float ly1 = fy - (float)y0, lx1 = fx - (float)x0;
  const float ly0 = 1.0f - ly1, lx0 = 1.0f - lx1;
  auto at = [&](int yy, int xx) {
    return *reinterpret_cast<const float4*>(b + (((size_t)n * h + yy) * w + xx) * C + c4 * 4);
  };
  const float4 v00 = at(y0, x0), v01 = at(y0, x1), v10 = at(y1, x0), v11 = at(y1, x1);
  const float4 av = *reinterpret_cast<const float4*>(a + i * 4);
  float4 o;
  o.x = av.x + (ly0 * (lx0 * v00.x + lx1 * v01.x) + ly1 * (lx0 * v10.x + lx1 * v11.x));
  o.y = av.y + (ly0 * (lx0 * v00.y + lx1 * v01.y) + ly1 * (lx0 * v10.y + lx1 * v11.y));
  o.z = av.z + (ly0 * (lx0 * v00.z + lx1 * v01.z) + ly1 * (lx0 * v10.z + lx1 * v11.z));
  o.w = av.w + (ly0 * (lx0 * v00.w + lx1 * v01.w) + ly1 * (lx0 * v10.w + lx1 * v11.w));
  *reinterpret_cast<float4*>(out + i * 4) = o;
}

// ------------------------------------------------------------------------------------------
// end_conv: ngf->2, 3x3 zero-padded, then / sigmas[y] (ncsnv2.py:512-516), fp32 throughout, fused with the final
// normalisation
// ------------------------------------------------------------------------------------------
// Fused tail of the network: normalizer (InstanceNorm++) -> ELU -> end_conv -> / sigmas[y] (ncsnv2.py:509-516)
// reading the raw fp32 trunk once, instead of materialising the activated tensor (a 268 MB write and a 3x re-read).
// One warp owns an 8-pixel-wide, kEndRows-tall strip and walks DOWN its input rows; a lane owns 4 input channels.
//  * every input row (10 columns x 16 bytes per lane) is prefetched with cp.async into a per-thread shared-memory
//    slot one row ahead, so the load latency hides behind the previous row's arithmetic;
//  * the row is normalised and activated once and feeds the three output rows it touches: accumulators r0 / r1 / r2
//    hold output rows hi+1 / hi / hi-1 and rotate after every input row (tap rows are added in the order 0, 1, 2);
//  * the 16 per-lane partial sums of a finished output row are reduced across the 32 lanes with a transposing
//    butterfly (8+4+2+1+1 shuffles instead of 16 x 5): lane l ends up with the total of output (co = l>>4, px = (l>>1)&7).
// Out-of-image inputs are the zero padding of the activated tensor.
constexpr int kEndRows = 16;
constexpr int kEndThreads = 128;
constexpr int kEndSmemBytes = 2 * 9 * 128 * 4 + 2 * 10 * kEndThreads * 16;    // weights + two row slots per thread

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NGF>
__global__ void __launch_bounds__(kEndThreads, 4)
end_conv_norm_kernel(const float* __restrict__ raw, const float* __restrict__ coef, const float* __restrict__ wgt,
                     const float* __restrict__ bias, const float* __restrict__ sigmas,
                     const int64_t* __restrict__ labels, float* __restrict__ out, int N, int H, int W, int fast_elu) {
  pdl_sync();
  static_assert(NGF == 128, "one float4 per lane");
  extern __shared__ __align__(16) uint8_t end_smem[];
  float (*sw)[9][NGF] = reinterpret_cast<float (*)[9][NGF]>(end_smem);                       // [2][9][NGF]
  float4* slots = reinterpret_cast<float4*>(end_smem + 2 * 9 * NGF * 4);                     // [2][10][kEndThreads]
  for (int i = threadIdx.x; i < 2 * 9 * NGF; i += blockDim.x) {
    const int co = i / (9 * NGF), r = i % (9 * NGF), tap = r / NGF, ci = r % NGF;
    sw[co][tap][ci] = wgt[((size_t)co * NGF + ci) * 9 + tap];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t strip = (size_t)blockIdx.x * (kEndThreads >> 5) + (threadIdx.x >> 5);
  const int W8 = W / 8, HB = H / kEndRows;
  if (strip >= (size_t)N * HB * W8) return;
  const int w0 = (int)(strip % W8) * 8, hb0 = (int)((strip / W8) % HB) * kEndRows, n = (int)(strip / ((size_t)W8 * HB));
  // this lane's 4 channels: {mean, a, b} triples are 12 contiguous floats
  float mu[4], ga[4], be[4];
  {
    const float4* cf4 = reinterpret_cast<const float4*>(coef + ((size_t)n * NGF + lane * 4) * 3);
    const float4 t0 = cf4[0], t1 = cf4[1], t2 = cf4[2];
    mu[0] = t0.x; ga[0] = t0.y; be[0] = t0.z; mu[1] = t0.w; ga[1] = t1.x; be[1] = t1.y;
    mu[2] = t1.z; ga[2] = t1.w; be[2] = t2.x; mu[3] = t2.y; ga[3] = t2.z; be[3] = t2.w;
  }
  const float sg = sigmas[labels[n]];
  const int co_l = lane >> 4, px_l = (lane >> 1) & 7;           // the output this lane holds after the reduction
  const float bias_l = bias[co_l];
  const int jlo = (w0 == 0) ? 1 : 0, jhi = (w0 + 8 == W) ? 9 : 10;   // in-image columns of the 10-column window

  auto prefetch = [&](int st) {                                 // input row hb0 - 1 + st into slot st & 1
    const int hi = hb0 - 1 + st;
    if (st < kEndRows + 2 && hi >= 0 && hi < H) {
      const float* row = raw + (((size_t)n * H + hi) * W + w0 - 1) * NGF + lane * 4;
      float4* dst = slots + (size_t)(st & 1) * 10 * kEndThreads + threadIdx.x;
#pragma unroll
      for (int j = 0; j < 10; ++j)
        if (j >= jlo && j < jhi) cp_async16(dst + j * kEndThreads, row + (size_t)j * NGF);
    }
    cp_async_commit();
  };
  auto act = [&](float v, int k) {
    const float t = ga[k] * (v - mu[k]) + be[k];
    return fast_elu ? elu_fast(t) : elu1(t);
  };

  float r0[2][8], r1[2][8], r2[2][8];                           // partial sums of output rows hi+1, hi, hi-1
#pragma unroll
  for (int i = 0; i < 8; ++i) r0[0][i] = r0[1][i] = r1[0][i] = r1[1][i] = r2[0][i] = r2[1][i] = 0.0f;
  prefetch(0);
#pragma unroll 1
  for (int st = 0; st < kEndRows + 2; ++st) {
    const int hi = hb0 - 1 + st;
    prefetch(st + 1);
    cp_async_wait<1>();                                         // row st has landed (a thread reads back only its own bytes)
    if (hi >= 0 && hi < H) {
      const float4* src = slots + (size_t)(st & 1) * 10 * kEndThreads + threadIdx.x;
      float4 col[10];
#pragma unroll
      for (int j = 0; j < 10; ++j) {
        col[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (j >= jlo && j < jhi) {
          const float4 v = src[j * kEndThreads];
          col[j] = make_float4(act(v.x, 0), act(v.y, 1), act(v.z, 2), act(v.w, 3));
        }
      }
      auto taps = [&](float (&acc)[2][8], int kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          const float4 w0v = *reinterpret_cast<const float4*>(&sw[0][kh * 3 + kw][lane * 4]);
          const float4 w1v = *reinterpret_cast<const float4*>(&sw[1][kh * 3 + kw][lane * 4]);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 v = col[i + kw];
            acc[0][i] = fmaf(v.w, w0v.w, fmaf(v.z, w0v.z, fmaf(v.y, w0v.y, fmaf(v.x, w0v.x, acc[0][i]))));
            acc[1][i] = fmaf(v.w, w1v.w, fmaf(v.z, w1v.z, fmaf(v.y, w1v.y, fmaf(v.x, w1v.x, acc[1][i]))));
          }
        }
      };
      if (hi + 1 < hb0 + kEndRows) taps(r0, 0);                 // rows outside [hb0, hb0 + kEndRows) belong to other strips
      if (hi >= hb0 && hi < hb0 + kEndRows) taps(r1, 1);
      if (hi - 1 >= hb0) taps(r2, 2);
    }
    const int ho = hi - 1;                                      // r2 is complete
    if (ho >= hb0 && ho < hb0 + kEndRows) {
      // transposing butterfly: at distance o a lane keeps the half of its values selected by its bit o and adds the
      // partner's copy of that half, so the value count halves per level
      float b[8], c[4], d[2], e;
      const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
      for (int k = 0; k < 8; ++k) {                             // value index = co * 8 + px: bit 4 of the lane selects co
        const float keep = h16 ? r2[1][k] : r2[0][k], send = h16 ? r2[0][k] : r2[1][k];
        b[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float keep = h8 ? b[k + 4] : b[k], send = h8 ? b[k] : b[k + 4];
        c[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float keep = h4 ? c[k + 2] : c[k], send = h4 ? c[k] : c[k + 2];
        d[k] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      {
        const float keep = h2 ? d[1] : d[0], send = h2 ? d[0] : d[1];
        e = keep + __shfl_xor_sync(0xffffffffu, send, 2);
      }
      e += __shfl_xor_sync(0xffffffffu, e, 1);
      if ((lane & 1) == 0) out[(((size_t)n * 2 + co_l) * H + ho) * W + w0 + px_l] = (e + bias_l) / sg;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      r2[0][i] = r1[0][i]; r2[1][i] = r1[1][i];
      r1[0][i] = r0[0][i]; r1[1][i] = r0[1][i];
      r0[0][i] = r0[1][i] = 0.0f;
    }
  }
  cp_async_wait<0>();
}

// NHWC fp32 -> NCHW fp32 (debug taps)
__global__ void nhwc_to_nchw_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int H, int W, int C) {
  const size_t total = (size_t)N * H * W * C;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % C);
  size_t t = i / C;
  const int w = (int)(t % W); t /= W;
  const int h = (int)(t % H);
  const int n = (int)(t / H);
  out[(((size_t)n * C + c) * H + h) * W + w] = in[i];
}

// ------------------------------------------------------------------------------------------
// weight repacking: torch [Cout][Cin][kh][kw] -> tensor-core [tap][Cout][Cin] (T) and
// SIMT [tap][Cin][Cout] (fp32)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ src, T* __restrict__ dst_tc, float* __restrict__ dst_simt,
                                   int Cout, int Cin, int taps, int tf32, T* __restrict__ dst_lo = nullptr) {
  const size_t total = (size_t)Cout * Cin * taps;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int tap = (int)(i % taps);
  const int ci = (int)((i / taps) % Cin);
  const int co = (int)(i / ((size_t)taps * Cin));
  const float v = src[i];
  if (dst_tc) {
    if constexpr (sizeof(T) == 2) {
      if (tf32) {                                             // fp16 arm: half bit patterns
        *reinterpret_cast<uint16_t*>(&dst_tc[((size_t)tap * Cout + co) * Cin + ci]) = pack1_h16(v, true);
      } else {
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        dst_tc[((size_t)tap * Cout + co) * Cin + ci] = hi;
        if (dst_lo) dst_lo[((size_t)tap * Cout + co) * Cin + ci] = __float2bfloat16_rn(v - __bfloat162float(hi));
      }
    } else {
      dst_tc[((size_t)tap * Cout + co) * Cin + ci] = tf32 ? round_tf32(v) : v;
    }
  }
  if (dst_simt) dst_simt[((size_t)tap * Cin + ci) * Cout + co] = v;
}

// ------------------------------------------------------------------------------------------
// fp32 CUDA-core implicit-GEMM convolution (strict-parity arm and validation of the tcgen05 path)
// tile: 64 pixels x 64 output channels, 256 threads, 4x4 outputs per thread, K step 16 channels
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
conv_simt_kernel(const float* __restrict__ in, const float* __restrict__ wgt /*[tap][Cin][Cout]*/, const ConvGeom g,
                 const EpiParams e) {
  pdl_sync();
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64];
  const int bw = g.W < 64 ? g.W : 64, bh = 64 / bw;
  const int tiles_w = g.W / bw, tiles_h = g.H / bh;
  const int tile = blockIdx.x;
  const int n = tile / (tiles_w * tiles_h);
  const int rem = tile % (tiles_w * tiles_h);
  const int h0 = (rem / tiles_w) * bh, w0 = (rem % tiles_w) * bw;
  const int co0 = blockIdx.y * 64;
  const int P = g.in_pad, Hp = g.H + 2 * P, Wp = g.W + 2 * P;
  const int t = threadIdx.x;
  const int ty = t / 16, tx = t % 16;
  // loader roles
  const int lp = t / 4, lk = (t % 4) * 4;            // A: pixel lp, channels lk..lk+3 of the 16-chunk
  const int lph = h0 + lp / bw, lpw = w0 + lp % bw;
  const int bk = t / 16, bj = (t % 16) * 4;          // B: k row bk, couts bj..bj+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int tap = 0; tap < g.taps; ++tap) {
    const int dy = (g.taps == 9) ? (tap / 3 - 1) * g.dil : 0;
    const int dx = (g.taps == 9) ? (tap % 3 - 1) * g.dil : 0;
    const float* arow = in + (((size_t)n * Hp + lph + P + dy) * Wp + lpw + P + dx) * g.Cin;
    const float* brow = wgt + (size_t)tap * g.Cin * g.Cout + co0;
    for (int k0 = 0; k0 < g.Cin; k0 += 16) {
      const float4 av = *reinterpret_cast<const float4*>(arow + k0 + lk);
      const float4 bv = *reinterpret_cast<const float4*>(brow + (size_t)(k0 + bk) * g.Cout + bj);
      __syncthreads();
      As[lk + 0][lp] = av.x; As[lk + 1][lp] = av.y; As[lk + 2][lp] = av.z; As[lk + 3][lp] = av.w;
      *reinterpret_cast<float4*>(&Bs[bk][bj]) = bv;
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
        const float aa[4] = {a.x, a.y, a.z, a.w}, bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(aa[i], bb[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = ty * 4 + i;
    epi_store<float, 4>(e, g, n, h0 + m / bw, w0 + m % bw, co0 + tx * 4, acc[i]);
  }
}

}  // namespace sdpc
