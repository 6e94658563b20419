// Shared types of the score-network kernels: convolution geometry, fused epilogue.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdpc {

// All activations are NHWC.  "raw" tensors are fp32 without halo; "operand" tensors carry a
// halo of `pad` pixels on every side of H and W (circular wrap or zeros, decided by their
// producer) so that a shifted 3x3 tap is always an in-bounds box for TMA.
struct ConvGeom {
  int N, H, W;        // output (= input) spatial size, batch
  int Cin, Cout;
  int taps;           // 9 (3x3) or 1 (1x1)
  int dil;            // dilation of the 3x3
  int in_pad;         // halo of the input operand (>= dil for 3x3)
  int BW, BH;         // pixel tile: BH rows x BW columns (BW a power of two, BW*BH = pixels per tile)
  int bw_shift;       // log2(BW)
  int tiles_w, tiles_h, num_tiles;
  int passes;         // 1, or 3 for the bf16x3 arm: (A_hi,B_hi), (A_hi,B_lo), (A_lo,B_hi) accumulated in TMEM
};

// Fused epilogue of every convolution (SIMT and tcgen05 kernels share it).
struct EpiParams {
  const float* bias;      // [Cout] or null
  const float* residual;  // fp32 raw [N,H,W,Cout] added after bias, or null
  float* out_raw;         // raw layout (no halo), value after bias + residual, or null; fp32 unless raw_bf16
  int raw_bf16;           // 1: out_raw holds bf16 (tensors only a normalisation pass reads; statistics stay fp32)
  void* out_acc;          // raw layout (no halo), value after bias only (CRP path: only a max-pool reads it), or null;
  int acc_bf16;           // 1: out_acc holds bf16 (max-pooling commutes with the monotone rounding), else fp32
  void* out_op;           // operand (T) with halo `op_pad`, value = act(after bias+residual), or null
  int op_pad;
  int op_elu;             // 1: ELU before the operand store
  int op_tf32;            // 1: the arm's alternate operand format - a 32-bit operand is rounded to tf32 (rna), a 16-bit operand
                          //    (and the 16-bit out_raw / out_acc copies) holds IEEE half instead of bf16 (fp16 arm)
  size_t op_lo_off;       // bf16x3 arm: element offset of the residual (lo) plane of out_op, else 0
  int prefetch_residual;  // 1: the epilogue warps prefetch the tile's residual rows into L2 before the accumulator is ready
#ifdef SDPC_DEV_HOOKS
  int dev_wrap;           // timing probe: bit 0 wraps store offsets, bit 1 wraps residual-load offsets into a 1 MB window
#endif
  float* stats;           // per-tile partial sums [tile][parts][Cout][2] (sum, sum of squares) of the out_raw
                          // values for InstanceNorm++ (parts = conv_umma_stats_parts(): 2 chunk parities per 256-pixel tile with swapped
                          // operands, 4 pixel quadrants per 128-pixel tile otherwise), or null
};

__device__ __forceinline__ float elu1(float v) { return v > 0.0f ? v : expm1f(v); }
// ex2.approx based ELU for the reduced-precision arms (abs error ~1e-7, far below bf16/tf32 rounding): one multiply,
// one MUFU, one add and a select (__expf would add denormal-range scaling around the MUFU).  The fp32 arm keeps expm1f.
__device__ __forceinline__ float elu_fast(float v) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * 1.4426950408889634f));
  return v > 0.0f ? v : e - 1.0f;
}
template <typename T>
__device__ __forceinline__ float elu_sel(float v, bool reduced) {
  if (sizeof(T) == 2 || reduced) return elu_fast(v);
  return elu1(v);
}

__device__ __forceinline__ float round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// 16-bit containers (declared __nv_bfloat16 throughout) hold bf16 bit patterns or, in the fp16 arm, IEEE half bit patterns:
// the `tf32` / `alt` flag that rounds a 32-bit operand to tf32 selects the half format for a 16-bit one ("the arm's
// alternate operand format").  Half conversions saturate finite values at +-65504 (NaN stays NaN).
__device__ __forceinline__ float sat_half(float v) { return fabsf(v) > 65504.0f ? copysignf(65504.0f, v) : v; }
__device__ __forceinline__ uint32_t pack2_h16(float a, float b, bool alt) {
  if (alt) {
    const __half2 t = __floats2half2_rn(sat_half(a), sat_half(b));
    return *reinterpret_cast<const uint32_t*>(&t);
  }
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack2_h16(uint32_t w, bool alt) {
  if (alt) return __half22float2(*reinterpret_cast<const __half2*>(&w));
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ uint16_t pack1_h16(float v, bool alt) {
  if (alt) { const __half t = __float2half_rn(sat_half(v)); return *reinterpret_cast<const uint16_t*>(&t); }
  const __nv_bfloat16 t = __float2bfloat16_rn(v);
  return *reinterpret_cast<const uint16_t*>(&t);
}

// lo_off != 0 (bf16x3 arm): the operand is stored as two bf16 planes, hi = bf16(v) at dst and the rounding
// residual lo = bf16(v - hi) at dst + lo_off, so that hi + lo carries ~16 mantissa bits.
template <typename T>
__device__ __forceinline__ void store_op4(T* dst, const float* v, bool tf32, size_t lo_off = 0);
template <>
__device__ __forceinline__ void store_op4<float>(float* dst, const float* v, bool tf32, size_t) {
  float4 o = tf32 ? make_float4(round_tf32(v[0]), round_tf32(v[1]), round_tf32(v[2]), round_tf32(v[3]))
                  : make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(dst) = o;
}
template <>
__device__ __forceinline__ void store_op4<__nv_bfloat16>(__nv_bfloat16* dst, const float* v, bool alt, size_t lo_off) {
  if (alt) {                                                  // fp16 arm: half bit patterns, single plane
    *reinterpret_cast<uint2*>(dst) = make_uint2(pack2_h16(v[0], v[1], true), pack2_h16(v[2], v[3], true));
    return;
  }
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 o;
  o.x = *reinterpret_cast<uint32_t*>(&a);
  o.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(dst) = o;
  if (lo_off) {
    const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    __nv_bfloat162 la = __floats2bfloat162_rn(v[0] - fa.x, v[1] - fa.y);
    __nv_bfloat162 lb = __floats2bfloat162_rn(v[2] - fb.x, v[3] - fb.y);
    uint2 l;
    l.x = *reinterpret_cast<uint32_t*>(&la);
    l.y = *reinterpret_cast<uint32_t*>(&lb);
    *reinterpret_cast<uint2*>(dst + lo_off) = l;
  }
}

// one element of an operand tensor (the swapped-operand epilogue stores a channel per lane)
template <typename T>
__device__ __forceinline__ void store_op1(T* dst, float v, bool tf32, size_t lo_off);
template <>
__device__ __forceinline__ void store_op1<float>(float* dst, float v, bool tf32, size_t) { *dst = tf32 ? round_tf32(v) : v; }
template <>
__device__ __forceinline__ void store_op1<__nv_bfloat16>(__nv_bfloat16* dst, float v, bool alt, size_t lo_off) {
  if (alt) { *reinterpret_cast<uint16_t*>(dst) = pack1_h16(v, true); return; }
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  *dst = hi;
  if (lo_off) dst[lo_off] = __float2bfloat16_rn(v - __bfloat162float(hi));
}

// Positions a pixel (h,w) of an HxW image occupies in a tensor padded by P with circular wrap: its
// interior position (h+P, w+P) plus at most one duplicate row and one duplicate column on the opposite
// border (H, W >= 2P, so a pixel is never near both borders of an axis).  hb / wb are -1 when absent.
struct HaloPos { int ha, hb, wa, wb; };
__device__ __forceinline__ HaloPos halo_pos(int h, int w, int H, int W, int P) {
  HaloPos d;
  d.ha = h + P;
  d.wa = w + P;
  d.hb = (h < P) ? h + P + H : ((h >= H - P) ? h + P - H : -1);
  d.wb = (w < P) ? w + P + W : ((w >= W - P) ? w + P - W : -1);
  return d;
}
// f(hp, wp) for every position the pixel owns
template <typename F>
__device__ __forceinline__ void for_each_halo_pos(const HaloPos& d, F f) {
  f(d.ha, d.wa);
  if (d.wb >= 0) f(d.ha, d.wb);
  if (d.hb >= 0) {
    f(d.hb, d.wa);
    if (d.wb >= 0) f(d.hb, d.wb);
  }
}

// Store NV (multiple of 4) consecutive output channels [c0, c0+NV) of pixel (n,h,w).
template <typename T, int NV>
__device__ __forceinline__ void epi_store(const EpiParams& e, const ConvGeom& g, int n, int h, int w, int c0,
                                          float* v) {
  const size_t pix = ((size_t)n * g.H + h) * g.W + w;
  if (e.bias) {
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(e.bias + c0 + i);
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (e.out_acc) {
    float* d = reinterpret_cast<float*>(e.out_acc) + pix * g.Cout + c0;       // SIMT path: always fp32
#pragma unroll
    for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (e.residual) {
    const float* r = e.residual + pix * g.Cout + c0;
#pragma unroll
    for (int i = 0; i < NV; i += 4) {
      float4 b = *reinterpret_cast<const float4*>(r + i);
      v[i] += b.x; v[i + 1] += b.y; v[i + 2] += b.z; v[i + 3] += b.w;
    }
  }
  if (e.out_raw) {
    float* d = e.out_raw + pix * g.Cout + c0;
#pragma unroll
    for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(d + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
  }
  if (e.out_op) {
    if (e.op_elu) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = elu_sel<T>(v[i], e.op_tf32 != 0);
    }
    const int P = e.op_pad, Hp = g.H + 2 * P, Wp = g.W + 2 * P;
    const HaloPos d = halo_pos(h, w, g.H, g.W, P);
    for_each_halo_pos(d, [&](int hp, int wp) {
      T* dst = reinterpret_cast<T*>(e.out_op) + (((size_t)n * Hp + hp) * Wp + wp) * g.Cout + c0;
#pragma unroll
      for (int i = 0; i < NV; i += 4) store_op4<T>(dst + i, v + i, e.op_tf32 != 0, e.op_lo_off);
    });
  }
}

}  // namespace sdpc
