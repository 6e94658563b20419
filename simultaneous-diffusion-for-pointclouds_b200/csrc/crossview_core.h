// Per-point / per-pixel arithmetic of the cross-view step, shared by the CUDA kernels
// (crossview.cu) and by the serial host emulation used in CPU tests (tests/host_emul).
//
// Operation order follows the reference op by op, because torch evaluates each op as its own
// kernel with one rounding per op: no FMA contraction is allowed here (crossview.cu is compiled
// with -fmad=false, the host build with -ffp-contract=off); the only fused multiply-adds are the
// explicit fma() chains that reproduce the 4-term dot products of torch.bmm
// (KITTISampling.py:185,205).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define SDPC_HD __host__ __device__ __forceinline__
#else
#define SDPC_HD inline
#endif

namespace sdpc {

struct GeoConsts {
  double h_min, dh, big_row_min, dv;
  int H, W, R;
  // torch divides a tensor by a Python/NumPy scalar differently per device: the CPU kernels divide,
  // the CUDA kernel multiplies by the reciprocal (BinaryDivTrueKernel.cu, "may lose one bit").
  // recip = 1 reproduces the CUDA reference bit-for-bit, recip = 0 the CPU reference.
  int recip;
};

SDPC_HD float sdiv(float a, float b, int recip) { return recip ? a * (1.0f / b) : a / b; }
SDPC_HD double sdiv(double a, double b, int recip) { return recip ? a * (1.0 / b) : a / b; }

// KITTISampling.py:161-166: realDistance = (2^(|x0|*6/sigmaMod) - 1) * sign, all float32.
SDPC_HD float decode_range(float x0, float sigma_mod, int recip) {
  float e = fabsf(x0) * 6.0f;
  e = sdiv(e, sigma_mod, recip);
  float d = powf(2.0f, e) - 1.0f;
  return (x0 < 0.0f) ? d * -1.0f : d * 1.0f;
}

// KITTISampling.py:176-178: float32 range promoted to float64 by the LUT multiply.
SDPC_HD void unproject(float dist, double ca, double sa, double ce, double se, double P[3]) {
  double d = (double)dist;
  P[0] = (d * ca) * ce;
  P[1] = (d * sa) * ce;
  P[2] = d * se;
}

// One row of a [4x4].[x,y,z,w] product in torch.bmm's float64 order: products accumulated in
// k order with fused multiply-adds (verified against torch CPU bmm bit-for-bit).
SDPC_HD double dot4(const double* m, double x, double y, double z, double w) {
  double acc = m[0] * x;
  acc = fma(m[1], y, acc);
  acc = fma(m[2], z, acc);
  acc = fma(m[3], w, acc);
  return acc;
}

struct Candidate {
  double nd;   // log-range of the point seen from the target view
  int row;     // row in the R-row grid (already flipped), may be out of range
  int col;
};

// KITTISampling.py:209-251: norm, log2, two atan2, round-half-even, flips.
SDPC_HD Candidate reproject(double qx, double qy, double qz, float sigma_mod, const GeoConsts& g) {
  Candidate c;
  double xy = qx * qx + qy * qy;
  double r = sqrt(xy + qz * qz);
  double nd = log2(r + 1.0);
  nd = sdiv(nd, 6.0, g.recip);
  c.nd = nd * (double)sigma_mod;
  double horiz = atan2(qy, qx);
  double vert = atan2(qz, sqrt(xy));
  double cf = rint(sdiv(horiz - g.h_min, g.dh, g.recip));
  double rf = rint(sdiv(vert - g.big_row_min, g.dv, g.recip));
  // .int() of an already rounded double; NaN/inf map to INT_MIN like x86 cvttsd2si
  int ci = (cf >= -2147483648.0 && cf <= 2147483647.0) ? (int)cf : INT32_MIN;
  int ri = (rf >= -2147483648.0 && rf <= 2147483647.0) ? (int)rf : INT32_MIN;
  c.col = (int)((unsigned)ci * (unsigned)-1 + (unsigned)(g.W - 1));   // int32 wrap-around like torch
  c.row = (int)((unsigned)ri * (unsigned)-1 + (unsigned)(g.R - 1));
  return c;
}

// Same result as reproject(), cheaper: the log-range (the z-buffer key) is always the exact float64 expression,
// but each of the two float64 atan2 calls is replaced by a float32 estimate whenever that estimate is provably
// far from a rounding boundary.  Error budget of the estimate: inputs rounded to fp32 (6e-8 relative each),
// atan2f <= 2 ulp at pi (5e-7 rad), i.e. < 2e-4 pixel; the guard band is 1e-3 pixel (5x the budget).  Points inside
// the guard band (about 0.2 % per axis), non-finite or huge coordinates take the float64 path, so the integers are
// identical.  The band is kept narrow because the float64 path is paid per WARP: with a 1e-2 band about half of the
// warps had at least one lane in it (ncu: the two fallback lines were 28 % of the kernel's instructions).
constexpr double kFastGuard = 1e-3;
SDPC_HD Candidate reproject_fast(double qx, double qy, double qz, float sigma_mod, const GeoConsts& g) {
  Candidate c;
  const double xy = qx * qx + qy * qy;
  const double r = sqrt(xy + qz * qz);
  double nd = log2(r + 1.0);
  nd = sdiv(nd, 6.0, g.recip);
  c.nd = nd * (double)sigma_mod;
  const bool tame = r < 1e15;
  const float fx = (float)qx, fy = (float)qy, fz = (float)qz;
  double cf = sdiv((double)atan2f(fy, fx) - g.h_min, g.dh, g.recip);
  double rc = rint(cf);
  if (!(tame && fabs(cf - rc) < 0.5 - kFastGuard)) rc = rint(sdiv(atan2(qy, qx) - g.h_min, g.dh, g.recip));
  double rf = sdiv((double)atan2f(fz, sqrtf(fx * fx + fy * fy)) - g.big_row_min, g.dv, g.recip);
  double rr = rint(rf);
  if (!(tame && fabs(rf - rr) < 0.5 - kFastGuard)) rr = rint(sdiv(atan2(qz, sqrt(xy)) - g.big_row_min, g.dv, g.recip));
  int ci = (rc >= -2147483648.0 && rc <= 2147483647.0) ? (int)rc : INT32_MIN;
  int ri = (rr >= -2147483648.0 && rr <= 2147483647.0) ? (int)rr : INT32_MIN;
  c.col = (int)((unsigned)ci * (unsigned)-1 + (unsigned)(g.W - 1));
  c.row = (int)((unsigned)ri * (unsigned)-1 + (unsigned)(g.R - 1));
  return c;
}

SDPC_HD bool in_grid(const Candidate& c, const GeoConsts& g) {
  return c.col > -1 && c.col < g.W && c.row > -1 && c.row < g.R;
}

// Fixed-point accumulation makes the per-pixel sums order independent (deterministic atomics).
constexpr double kDepthScale = 1099511627776.0;      // 2^40
constexpr double kIntenScale = 4294967296.0;         // 2^32
SDPC_HD long long depth_to_fixed(double nd) { return (long long)rint(nd * kDepthScale); }
SDPC_HD long long inten_to_fixed(float v) { return (long long)rint((double)v * kIntenScale); }

struct Fused {
  double depth;  // float64 log-range of the shared image at this grid cell
  float inten;
  bool filled;
};

// KITTISampling.py:348-394 for one grid cell: average, optional controlled average, re-log.
SDPC_HD Fused fuse_cell(unsigned cnt, long long sum_d_fx, long long sum_i_fx, double min_d, float min_i,
                        float sigma_mod, double allowance, int recip) {
  Fused f;
  f.filled = cnt > 0;
  float scaling = (float)cnt + 0.000000001f;                  // float32, as in the reference
  double avg_d = ((double)sum_d_fx / kDepthScale) / (double)scaling;
  float avg_i = (float)((double)sum_i_fx / kIntenScale) / scaling;
  if (!f.filled) { min_d = 0.0; min_i = 0.0f; }
  if (allowance >= 0.0) {
    double sm = (double)sigma_mod;
    double m_avg = pow(2.0, sdiv(fabs(avg_d) * 6.0, sm, recip)) - 1.0;
    double m_min = pow(2.0, sdiv(fabs(min_d) * 6.0, sm, recip)) - 1.0;
    bool far = m_avg > m_min + allowance;
    if (far) { avg_i = min_i; m_avg = m_min + allowance / 5.0; }
    avg_d = sdiv(log2(m_avg + 1.0), 6.0, recip) * sm;
  }
  f.depth = avg_d;
  f.inten = avg_i;
  return f;
}

// torch.nan_to_num defaults (KITTISampling.py:138)
SDPC_HD float nan_to_num(float v) {
  if (v != v) return 0.0f;
  if (v > 3.4028234663852886e38f) return 3.4028234663852886e38f;
  if (v < -3.4028234663852886e38f) return -3.4028234663852886e38f;
  return v;
}

// KITTISampling.py:144,156, left to right in float32, one rounding per op.
SDPC_HD float langevin_value(float x, float g, float ref, int mask, float z, float eps, float rho,
                             float noise_scale, float* grad_likelihood) {
  float gl = (float)(-mask) * (x - ref);
  *grad_likelihood = gl;
  float a = x + eps * g;
  a = a + rho * gl;
  a = a + z * noise_scale;
  return a;
}

// tooHigh gate (KITTISampling.py:162): max|x0| * 6 / sigmaMod > 50 in float32
SDPC_HD bool too_high_gate(float max_abs, float sigma_mod, int recip) {
  return sdiv(max_abs * 6.0f, sigma_mod, recip) > 50.0f;
}

}  // namespace sdpc
