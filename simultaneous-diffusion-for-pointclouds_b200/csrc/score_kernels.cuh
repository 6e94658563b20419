// CUDA-core kernels of the score network: everything that is not a 128/256-channel convolution.
// (begin/end convolutions, InstanceNorm++ statistics, operand materialisation with halo,
// max / mean pooling, bilinear x2, layout conversion, fp32 SIMT convolution for the strict arm.)
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

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
__global__ void __launch_bounds__(256)
begin_conv_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                  float* __restrict__ out, int N, int H, int W) {
  constexpr int K = 36;
  __shared__ float sw[K][NGF];
  __shared__ float sin_[64][K + 1];
  const int tiles_w = W / 64;
  const int blk = blockIdx.x;
  const int w0 = (blk % tiles_w) * 64;
  const int h = (blk / tiles_w) % H;
  const int n = blk / (tiles_w * H);
  for (int i = threadIdx.x; i < K * NGF; i += blockDim.x) {
    const int co = i / K, k = i % K;                 // wgt is [co][ci][kh][kw] = [co][k]
    sw[k][co] = wgt[i];
  }
  for (int i = threadIdx.x; i < 64 * K; i += blockDim.x) {
    const int p = i / K, k = i % K;
    const int ci = k / 9, kh = (k % 9) / 3, kw = k % 3;
    const int hh = h + kh - 1, ww = w0 + p + kw - 1;
    float v = 0.0f;
    if (hh >= 0 && hh < H && ww >= 0 && ww < W) {
      if (ci < 2) v = 2.0f * x[(((size_t)n * 2 + ci) * H + hh) * W + ww] - 1.0f;
      else if (ci == 2) v = linspace01(ww, W);
      else v = linspace01(hh, H);
    }
    sin_[p][k] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * NGF; i += blockDim.x) {
    const int p = i / NGF, c = i % NGF;
    float acc = 0.0f;
#pragma unroll
    for (int k = 0; k < K; ++k) acc = fmaf(sin_[p][k], sw[k][c], acc);
    out[(((size_t)n * H + h) * W + w0 + p) * NGF + c] = acc + bias[c];
  }
}

// ------------------------------------------------------------------------------------------
// InstanceNorm++ (normalization.py:163-176)
// ------------------------------------------------------------------------------------------
// per-(n,c) sum and sum of squares over H*W; fp32 partials per thread, fp64 across threads.
__global__ void __launch_bounds__(256)
stats_kernel(const float* __restrict__ in, double* __restrict__ stats, int HW, int C, int pix_per_block) {
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
__global__ void norm_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ alpha,
                                     const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ coef, int HW, int C) {
  __shared__ double sm[256];
  __shared__ double s_m, s_v;
  const int n = blockIdx.x, c = threadIdx.x;
  const double S = stats[((size_t)n * C + c) * 2], Q = stats[((size_t)n * C + c) * 2 + 1];
  const double mean = S / HW;
  double var = Q / HW - mean * mean;
  if (var < 0.0) var = 0.0;
  sm[c] = mean;
  __syncthreads();
  if (c == 0) { double t = 0; for (int i = 0; i < C; ++i) t += sm[i]; s_m = t / C; }
  __syncthreads();
  const double dm = mean - s_m;
  sm[c] = dm * dm;
  __syncthreads();
  if (c == 0) { double t = 0; for (int i = 0; i < C; ++i) t += sm[i]; s_v = t / (C - 1); }
  __syncthreads();
  const float rstd = 1.0f / sqrtf((float)var + 1e-5f);
  const float mean_n = (float)dm / sqrtf((float)s_v + 1e-5f);
  float* o = coef + ((size_t)n * C + c) * 3;
  o[0] = (float)mean;
  o[1] = gamma[c] * rstd;
  o[2] = gamma[c] * (mean_n * alpha[c]) + beta[c];
}

// ------------------------------------------------------------------------------------------
// operand materialisation: raw fp32 NHWC -> T NHWC with halo
// ------------------------------------------------------------------------------------------
enum { OP_COPY = 0, OP_ELU = 1, OP_NORM_ELU = 2 };
enum { HALO_CIRC = 0, HALO_ZERO = 1 };

template <typename T>
__device__ __forceinline__ void store_op8(T* dst, const float* v, bool tf32) {
  store_op4<T>(dst, v, tf32);
  store_op4<T>(dst + 4, v + 4, tf32);
}

template <typename T>
__global__ void __launch_bounds__(256)
to_operand_kernel(const float* __restrict__ in, const float* __restrict__ coef, T* __restrict__ out, int N, int H,
                  int W, int C, int P, int mode, int halo, int tf32) {
  const int Hp = H + 2 * P, Wp = W + 2 * P, C8 = C / 8;
  const size_t total = (size_t)N * Hp * Wp * C8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  size_t t = i / C8;
  const int wp = (int)(t % Wp); t /= Wp;
  const int hp = (int)(t % Hp);
  const int n = (int)(t / Hp);
  int h = hp - P, w = wp - P;
  float v[8];
  const bool outside = h < 0 || h >= H || w < 0 || w >= W;
  if (outside && halo == HALO_ZERO) {
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = 0.0f;
  } else {
    h = (h % H + H) % H;
    w = (w % W + W) % W;
    const float* src = in + (((size_t)n * H + h) * W + w) * C + c8 * 8;
    const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    if (mode == OP_NORM_ELU) {
      const float* cf = coef + ((size_t)n * C + c8 * 8) * 3;
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = elu1(cf[k * 3 + 1] * (v[k] - cf[k * 3]) + cf[k * 3 + 2]);
    } else if (mode == OP_ELU) {
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] = elu1(v[k]);
    }
  }
  store_op8<T>(out + i * 8, v, tf32 != 0);
}

// CRP stage (layers.py:76-83): out_op = [ELU](maxpool5(in)) with circular halo for the conv that
// follows; MaxPool2d(5,1,2) itself pads with -inf, i.e. the window is clipped at the image border.
// ELU is monotone, so ELU(maxpool(x)) == maxpool(ELU(x)).  Optionally also emits x0 = ELU(in).
template <typename T>
__global__ void __launch_bounds__(256)
maxpool5_kernel(const float* __restrict__ in, float* __restrict__ x0_out, T* __restrict__ out, int N, int H, int W,
                int C, int P, int elu_in, int tf32) {
  const int Hp = H + 2 * P, Wp = W + 2 * P, C8 = C / 8;
  const size_t total = (size_t)N * Hp * Wp * C8;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c8 = (int)(i % C8);
  size_t t = i / C8;
  const int wp = (int)(t % Wp); t /= Wp;
  const int hp = (int)(t % Hp);
  const int n = (int)(t / Hp);
  const bool interior = hp >= P && hp < H + P && wp >= P && wp < W + P;
  const int h = ((hp - P) % H + H) % H, w = ((wp - P) % W + W) % W;
  float m[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  const int h_lo = max(h - 2, 0), h_hi = min(h + 2, H - 1), w_lo = max(w - 2, 0), w_hi = min(w + 2, W - 1);
  for (int hh = h_lo; hh <= h_hi; ++hh)
    for (int ww = w_lo; ww <= w_hi; ++ww) {
      const float* src = in + (((size_t)n * H + hh) * W + ww) * C + c8 * 8;
      const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
      m[0] = fmaxf(m[0], a.x); m[1] = fmaxf(m[1], a.y); m[2] = fmaxf(m[2], a.z); m[3] = fmaxf(m[3], a.w);
      m[4] = fmaxf(m[4], b.x); m[5] = fmaxf(m[5], b.y); m[6] = fmaxf(m[6], b.z); m[7] = fmaxf(m[7], b.w);
    }
  if (elu_in) {
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = elu1(m[k]);
  }
  store_op8<T>(out + i * 8, m, tf32 != 0);
  if (x0_out && interior) {
    const size_t o = (((size_t)n * H + h) * W + w) * C + c8 * 8;
    const float4 a = *reinterpret_cast<const float4*>(in + o), b = *reinterpret_cast<const float4*>(in + o + 4);
    float4 ea = make_float4(elu1(a.x), elu1(a.y), elu1(a.z), elu1(a.w));
    float4 eb = make_float4(elu1(b.x), elu1(b.y), elu1(b.z), elu1(b.w));
    *reinterpret_cast<float4*>(x0_out + o) = ea;
    *reinterpret_cast<float4*>(x0_out + o + 4) = eb;
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
                float* __restrict__ out_raw, int N, int H, int W, int C, int tf32) {
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
    store_op4<T>(out_op + i * 4, vv, tf32 != 0);
  }
}

// out = a + bilinear_x2(b), align_corners=True (layers.py:182; F.interpolate)
__global__ void __launch_bounds__(256)
upsample_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int N, int H,
                    int W, int C, int h, int w) {
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
// end_conv: ngf->2, 3x3 zero-padded, then / sigmas[y] (ncsnv2.py:512-516).  fp32 throughout.
// One warp per output pixel; the operand has a zero halo of 1.
// ------------------------------------------------------------------------------------------
template <int NGF>
__global__ void __launch_bounds__(256)
end_conv_kernel(const float* __restrict__ op, const float* __restrict__ wgt, const float* __restrict__ bias,
                const float* __restrict__ sigmas, const int64_t* __restrict__ labels, float* __restrict__ out, int N,
                int H, int W) {
  static_assert(NGF == 128, "one float4 per lane");
  __shared__ float sw[2][9][NGF];
  for (int i = threadIdx.x; i < 2 * 9 * NGF; i += blockDim.x) {
    const int co = i / (9 * NGF), r = i % (9 * NGF), tap = r / NGF, ci = r % NGF;
    sw[co][tap][ci] = wgt[((size_t)co * NGF + ci) * 9 + tap];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const size_t pix = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pix >= (size_t)N * H * W) return;
  const int w = (int)(pix % W), h = (int)((pix / W) % H), n = (int)(pix / ((size_t)W * H));
  const int Hp = H + 2, Wp = W + 2;
  float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const float4 v = *reinterpret_cast<const float4*>(
        op + (((size_t)n * Hp + h + tap / 3) * Wp + w + tap % 3) * NGF + lane * 4);
    const float4 w0 = *reinterpret_cast<const float4*>(&sw[0][tap][lane * 4]);
    const float4 w1 = *reinterpret_cast<const float4*>(&sw[1][tap][lane * 4]);
    a0 += v.x * w0.x + v.y * w0.y + v.z * w0.z + v.w * w0.w;
    a1 += v.x * w1.x + v.y * w1.y + v.z * w1.z + v.w * w1.w;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  }
  if (lane == 0) {
    const float s = sigmas[labels[n]];
    out[(((size_t)n * 2 + 0) * H + h) * W + w] = (a0 + bias[0]) / s;
    out[(((size_t)n * 2 + 1) * H + h) * W + w] = (a1 + bias[1]) / s;
  }
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
                                   int Cout, int Cin, int taps, int tf32) {
  const size_t total = (size_t)Cout * Cin * taps;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int tap = (int)(i % taps);
  const int ci = (int)((i / taps) % Cin);
  const int co = (int)(i / ((size_t)taps * Cin));
  const float v = src[i];
  if (dst_tc) {
    if constexpr (sizeof(T) == 2) dst_tc[((size_t)tap * Cout + co) * Cin + ci] = __float2bfloat16_rn(v);
    else dst_tc[((size_t)tap * Cout + co) * Cin + ci] = tf32 ? round_tf32(v) : v;
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
