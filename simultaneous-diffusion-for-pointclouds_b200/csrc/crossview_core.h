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
#include <string.h>

#include "log2_table.h"

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
  double inv_dh, inv_dv;   // 1.0 / dh, 1.0 / dv (IEEE division): the pixel estimates of pixel_fast()
};

inline GeoConsts make_geo(double h_min, double dh, double big_row_min, double dv, int H, int W, int R, int recip) {
  GeoConsts g;
  g.h_min = h_min; g.dh = dh; g.big_row_min = big_row_min; g.dv = dv;
  g.H = H; g.W = W; g.R = R; g.recip = recip ? 1 : 0;
  g.inv_dh = 1.0 / dh; g.inv_dv = 1.0 / dv;
  return g;
}

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

// ---- production path of the scatter: pixel indices first (cheap), the log-range only for candidates that land ----
// Squared range of a candidate in the reference's operation order (KITTISampling.py:209-212).  The log-range below is
// a monotone function of it, so the production z-buffer keeps min(r^2) (5 float64 operations per candidate, exact) and
// the nearest log-range is evaluated once per cell from the winning r^2 - the same float64 value the reference's
// per-candidate expression gives for that candidate, bit for bit.
SDPC_HD double range2(double qx, double qy, double qz) {
  const double xy = qx * qx + qy * qy;
  return xy + qz * qz;
}
// log2(sqrt(r2) + 1) / 6 * sigmaMod, the float64 expression of reproject()
SDPC_HD double log_range_of_r2(double r2, float sigma_mod, const GeoConsts& g) {
  double nd = log2(sqrt(r2) + 1.0);
  nd = sdiv(nd, 6.0, g.recip);
  return nd * (double)sigma_mod;
}
SDPC_HD double log_range(double qx, double qy, double qz, float sigma_mod, const GeoConsts& g) {
  return log_range_of_r2(range2(qx, qy, qz), sigma_mod, g);
}

// log2 for the depth SUM of the production scatter: exponent + 128-entry table on the top mantissa bits + degree-6
// polynomial of the residual (|t| <= 2^-8), 9 float64 operations instead of the ~50 instructions of the library call.
// v must be a normal float64 >= 1 (it is r + 1).  Error against the exact logarithm: <= 1 ulp of the result, 3.7e-15
// absolute for results up to 50 (tools/gen_log2_table.py checks 4e6 values) - the same accuracy class as the library's
// log2, but not the same bits, so it is used only where the reference itself carries no bit-exact meaning: the
// per-cell sum of log-ranges, which is accumulated in 2^-40 fixed point (9e-13) and ends as a float32.
// tab: the {rc, T} pairs of log2_table.h (shared memory on the device).
#define SDPC_LOG2_TABLE_DOUBLES 256
constexpr double kRoundMagicLog = 6755399441055744.0;          // 1.5 * 2^52
SDPC_HD double fast_log2(double v, const double* tab) {
  const double c[6] = {SDPC_LOG2_COEF};
  long long bits;
  memcpy(&bits, &v, 8);
  const int e = (int)(bits >> 52) - 1023;
  const int i = (int)(bits >> 45) & 127;
  const long long mb = (bits & 0x000FFFFFFFFFFFFFll) | 0x3FF0000000000000ll;
  double m;
  memcpy(&m, &mb, 8);
  const long long eb = 0x4338000000000000ll + (long long)e;       // (double)e without the conversion unit
  double ed;
  memcpy(&ed, &eb, 8);
  ed -= kRoundMagicLog;
  const double t = fma(m, tab[2 * i], -1.0);
  double p = c[5];
  p = fma(p, t, c[4]);
  p = fma(p, t, c[3]);
  p = fma(p, t, c[2]);
  p = fma(p, t, c[1]);
  p = fma(p, t, c[0]);
  return ed + fma(t, p, tab[2 * i + 1]);
}
SDPC_HD double fast_log_range_of_r2(double r2, float sigma_mod, const GeoConsts& g, const double* tab) {
  double nd = fast_log2(sqrt(r2) + 1.0, tab);
  nd = sdiv(nd, 6.0, g.recip);
  return nd * (double)sigma_mod;
}

// atan2 estimate for pixel indices only (never for a value that is returned): min/max ratio, degree-15 odd minimax
// polynomial (8 coefficients, fitted error 3.7e-8 rad), quadrant fix-ups.  Measured over 2.4e7 points incl. the axes and
// diagonals, with the ratio perturbed by +-2 ulp like the approximate division: |error| <= 4.7e-7 rad = 7.6e-5 column
// pixels / 6.1e-5 row pixels (the fit and the float32 emulation: tools/fit_fast_atan.py).  x = y = 0 gives NaN, which
// fails the guard test below and takes the exact path.
SDPC_HD float fast_atan2f(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
#if defined(__CUDA_ARCH__)
  const float t = __fdividef(mn, mx);
#else
  const float t = mn / mx;
#endif
  const float s = t * t;
  float p = -4.054529127e-03f;
  p = fmaf(p, s, 2.186281234e-02f);
  p = fmaf(p, s, -5.591210350e-02f);
  p = fmaf(p, s, 9.642180055e-02f);
  p = fmaf(p, s, -1.390862167e-01f);
  p = fmaf(p, s, 1.994656473e-01f);
  p = fmaf(p, s, -3.332985938e-01f);
  p = fmaf(p, s, 9.999993443e-01f);
  float a = t * p;
  if (ay > ax) a = 1.57079632679489662f - a;
  if (x < 0.0f) a = 3.14159265358979324f - a;
  return copysignf(a, y);
}

// round-half-to-even of a float64 with |v| < 2^31 without the conversion unit: adding 1.5 * 2^52 leaves the rounded
// integer in the low mantissa bits (two's complement in the low 32 bits), subtracting it again gives rint(v)
constexpr double kRoundMagic = 6755399441055744.0;
SDPC_HD int magic_low32(double biased) {
  long long b;
  memcpy(&b, &biased, 8);
  return (int)(unsigned)(b & 0xFFFFFFFFll);
}

// Same integers as reproject(), cheaper: each float64 atan2 is replaced by fast_atan2f whenever the estimated pixel
// coordinate is provably far from a rounding boundary.  Error budget of the estimate: fast_atan2f 7.6e-5 px, the
// float32 rounding of its inputs 1e-5 px, the approximate square root of the planar norm 2e-5 px: < 1.1e-4 px; guard
// band 4e-4 px (3.6x).  Candidates inside the band (0.08 % per axis), non-finite or huge coordinates take the float64
// expressions of the reference, so the integers are identical (tests/host_emul checks every candidate of every CPU
// test case; the GPU tests compare with the full float64 kernel).  Two stages so that a kernel can run the estimates
// of several candidates without a branch between them:
//   pixel_estimate : biased (+ kRoundMagic) column / row coordinates and whether both are safely inside a pixel
//   pixel_exact    : the reference's float64 expressions, biased the same way
//   pixel_finish   : validity and the flipped integers.  Validity is decided on the rounded doubles: col = W-1-rc is
//                    inside [0, W) iff rc is inside [0, W-1]; NaN fails every comparison (the reference's INT_MIN is
//                    outside the grid too).
constexpr double kPixelGuard = 4e-4;
SDPC_HD bool pixel_estimate(double qx, double qy, double qz, const GeoConsts& g, double* cm, double* rm) {
  const float fx = (float)qx, fy = (float)qy, fz = (float)qz;
  const bool tame = fmaxf(fmaxf(fabsf(fx), fabsf(fy)), fabsf(fz)) < 1e15f;      // false for NaN / inf / huge
  const float xy = fx * fx + fy * fy;
#if defined(__CUDA_ARCH__)
  const float rxy = xy * rsqrtf(xy);                                            // 0 * inf = NaN for xy = 0: exact path
#else
  const float rxy = sqrtf(xy);
#endif
  const double cf = ((double)fast_atan2f(fy, fx) - g.h_min) * g.inv_dh;
  const double rf = ((double)fast_atan2f(fz, rxy) - g.big_row_min) * g.inv_dv;
  *cm = cf + kRoundMagic;
  *rm = rf + kRoundMagic;
  return tame && fabs(cf - (*cm - kRoundMagic)) < 0.5 - kPixelGuard && fabs(rf - (*rm - kRoundMagic)) < 0.5 - kPixelGuard;
}
SDPC_HD void pixel_exact(double qx, double qy, double qz, const GeoConsts& g, double* cm, double* rm) {
  // rint() of the reference first: the biased sum of an already integral double is exact
  *cm = rint(sdiv(atan2(qy, qx) - g.h_min, g.dh, g.recip)) + kRoundMagic;
  *rm = rint(sdiv(atan2(qz, sqrt(qx * qx + qy * qy)) - g.big_row_min, g.dv, g.recip)) + kRoundMagic;
}
SDPC_HD bool pixel_finish(double cm, double rm, const GeoConsts& g, int* row, int* col) {
  const double rc = cm - kRoundMagic, rr = rm - kRoundMagic;
  const bool ok = rc >= 0.0 && rc <= (double)(g.W - 1) && rr >= 0.0 && rr <= (double)(g.R - 1);
  *col = g.W - 1 - magic_low32(cm);
  *row = g.R - 1 - magic_low32(rm);
  return ok;
}
SDPC_HD bool pixel_fast(double qx, double qy, double qz, const GeoConsts& g, int* row, int* col) {
  double cm, rm;
  if (!pixel_estimate(qx, qy, qz, g, &cm, &rm)) pixel_exact(qx, qy, qz, g, &cm, &rm);
  return pixel_finish(cm, rm, g, row, col);
}

// Fixed-point accumulation makes the per-pixel sums order independent (deterministic atomics).
constexpr double kDepthScale = 1099511627776.0;      // 2^40
constexpr double kIntenScale = 4294967296.0;         // 2^32
SDPC_HD long long depth_to_fixed(double nd) { return (long long)rint(nd * kDepthScale); }
// the same value for 0 <= nd < 2048 (it is: nd <= log2(1e15 + 1) / 6 * 50) without the conversion unit
SDPC_HD long long depth_to_fixed_magic(double nd) {
  const double biased = nd * kDepthScale + kRoundMagic;
  long long b;
  memcpy(&b, &biased, 8);
  return b - 0x4338000000000000ll;
}
// Non-finite intensities enter the sum as 0 (the reference's float sum would make the cell NaN), finite ones are clamped to
// +-2^20 so that 2^11 candidates of one cell cannot overflow the 64-bit sum (deviations).
SDPC_HD long long inten_to_fixed(float v) {
  if (!(fabsf(v) <= 3.4028234663852886e38f)) return 0;
  const float c = fminf(fmaxf(v, -1048576.0f), 1048576.0f);
  return (long long)rint((double)c * kIntenScale);
}

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

// Production fusion of one grid cell from the z-buffer's min(r^2).  Same decisions as fuse_cell(); what it skips:
//  * a cell with one candidate is its own average and nearest depth: never "far";
//  * the far test (m_avg > m_min + allowance, metres) is first made on float32 estimates: m_avg + 1 = exp2f(.) of the
//    average, m_min + 1 = sqrtf(r2min) + 1 (the reference's 2^(6 min_d / sigmaMod) undoes the logarithm of the nearest
//    range up to 1e-14 relative).  Their relative error is below 1.6e-6 (input rounding 1.3e-6 for log2(r+1) < 64,
//    exp2f 2 ulp), the guard is 4e-6 of the two magnitudes, and the reference's float64 pow expressions decide inside
//    the guard or when an estimate is not finite;
//  * a cell that is not far keeps its average: the reference sends it through pow(2, .) - 1 and log2(. + 1) again, a
//    float64 round trip that moves the value by a few 1e-16 relative - its float32 cast (what newImages holds) differs
//    from the cast of the average in about 1e-8 of the cells, by one float32 ulp (test tolerance of newImages: 1e-5).
// far: the cell takes the nearest candidate's intensity and nearest depth + allowance / 5 (fuse_far, exact arithmetic).
struct FusedFast {
  double depth;
  float inten;
  bool filled, far;
};
SDPC_HD FusedFast fuse_cell_fast(unsigned cnt, long long sum_d_fx, long long sum_i_fx, double r2min, float sigma_mod,
                                 double allowance, const GeoConsts& g) {
  FusedFast f;
  f.filled = cnt > 0;
  f.far = false;
  const float scaling = (float)cnt + 0.000000001f;
  f.depth = ((double)sum_d_fx / kDepthScale) / (double)scaling;
  f.inten = (float)((double)sum_i_fx / kIntenScale) / scaling;
  if (allowance < 0.0 || cnt < 2) return f;
  const double sm = (double)sigma_mod;
  const double xa = sdiv(fabs(f.depth) * 6.0, sm, g.recip);
  const float ea = exp2f((float)xa), em = sqrtf((float)r2min) + 1.0f;
  const float diff = (ea - em) - (float)allowance;
  if (fabsf(diff) > 4e-6f * (ea + em)) {
    f.far = diff > 0.0f;
  } else {
    const double xm = sdiv(fabs(log_range_of_r2(r2min, sigma_mod, g)) * 6.0, sm, g.recip);
    f.far = (pow(2.0, xa) - 1.0) > (pow(2.0, xm) - 1.0) + allowance;
  }
  return f;
}
// depth and intensity of a far cell (KITTISampling.py:381-394): nearest candidate + allowance / 5, its intensity
SDPC_HD void fuse_far(FusedFast* f, double min_d, float min_i, float sigma_mod, double allowance, int recip) {
  const double sm = (double)sigma_mod;
  const double m_min = pow(2.0, sdiv(fabs(min_d) * 6.0, sm, recip)) - 1.0;
  const double m_avg = m_min + allowance / 5.0;
  f->depth = sdiv(log2(m_avg + 1.0), 6.0, recip) * sm;
  f->inten = min_i;
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
