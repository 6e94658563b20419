// TEST INFRASTRUCTURE: serial host emulation of crossview.cu's kernels, built with g++ from the
// same per-element header (csrc/crossview_core.h).  Lets the CPU test-suite check the kernels'
// arithmetic and indexing against the oracle before any GPU time is spent.  Never shipped.
#include <algorithm>
#include <cstring>
#include <limits>
#include <vector>

#include "../../include/sdpc_b200.h"
#include "../../simultaneous-diffusion-for-pointclouds_b200/csrc/crossview_core.h"

using namespace sdpc;

// The emulation exports the product library's entry-point names (host pointers instead of device
// pointers, `stream` ignored) so that StepRunner / ViewShard can be driven unchanged in CPU tests.
extern "C" size_t sdpc_step_workspace_bytes(int, int, int, int) { return 256; }
extern "C" int sdpc_step_workspace_init(void*, size_t, int, int, int, int, void*) { return 0; }
extern "C" const char* sdpc_last_error(void) { return "host emulation"; }

extern "C" int sdpc_langevin_update(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace, size_t,
                                    void*) {
  const int B = p->n_views, H = p->height, W = p->width;
  const int HW = H * W;
  const int t0 = p->tgt_first, tn = p->tgt_count ? p->tgt_count : B;
  float mx = 0.0f;
  bool has_nan = false;
  for (int v = t0; v < t0 + tn; ++v)
    for (int ch = 0; ch < 2; ++ch)
      for (int q = 0; q < HW; ++q) {
        size_t e = ((size_t)v * 2 + ch) * HW + q;
        float g = b->grad ? b->grad[e] : 0.0f;
        if (p->nan_to_num) g = nan_to_num(g);
        float z = b->noise ? b->noise[e] : 0.0f;
        float gl;
        float o = langevin_value(b->x[e], g, b->refer[e], b->mask[e], z, p->step_size, p->grad_ref, p->noise_scale, &gl);
        b->x[e] = o;
        if (b->grad_likelihood) b->grad_likelihood[e] = gl;
        if (ch == 0) { if (o != o) has_nan = true; mx = std::max(mx, fabsf(o)); }
      }
  if (has_nan) mx = std::numeric_limits<float>::quiet_NaN();
  *(float*)workspace = mx;
  return 0;
}

extern "C" int sdpc_step_read_max(void* workspace, float* out_max, void*) { *out_max = *(float*)workspace; return 0; }
extern "C" int sdpc_step_merge_max(void* workspace, const float* other, int n, void*) {
  float m = *(float*)workspace;
  for (int i = 0; i < n; ++i) { if (other[i] != other[i] || m != m) m = std::numeric_limits<float>::quiet_NaN(); else m = std::max(m, fabsf(other[i])); }
  *(float*)workspace = m;
  return 0;
}

extern "C" int sdpc_crossview_share(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace, size_t,
                                    void*) {
  const int B = p->n_views, A = p->group_size, H = p->height, W = p->width, R = p->big_rows;
  const int HW = H * W;
  const int t0 = p->tgt_first, tn = p->tgt_count ? p->tgt_count : B;
  const float mx = *(float*)workspace;
  const GeoConsts geo = make_geo(p->h_min, p->dh, p->big_row_min, p->dv, H, W, R, p->scalar_div_recip ? 1 : 0);
  const size_t cells = (size_t)B * R * W;
  std::vector<unsigned long long> zmin(cells, ~0ull);
  std::vector<unsigned> winner(cells, ~0u), cnt(cells, 0u);
  std::vector<long long> sum_d(cells, 0), sum_i(cells, 0);
  std::vector<double> r2min(cells, std::numeric_limits<double>::infinity());
  static const double log2_tab[SDPC_LOG2_TABLE_DOUBLES] = {SDPC_LOG2_TABLE};
  long long fast_mismatch = 0;
  for (int pass = 0; pass < 2; ++pass)
    for (int g = 0; g < B / A; ++g)
      for (int sa = 0; sa < A; ++sa) {
        const int bsrc = g * A + sa;
        const int t_lo = std::max(g * A, t0), t_hi = std::min((g + 1) * A, t0 + tn);
        for (int q = 0; q < HW; ++q) {
          bool src_ok = b->exist[(size_t)sa * HW + q] != 0;
          if (p->sky_filter) src_ok = src_ok && b->sky[(size_t)bsrc * HW + q] != 0;
          const int r = q / W, c = q % W;
          const float x0 = b->x[((size_t)bsrc * 2) * HW + q], x1 = b->x[((size_t)bsrc * 2 + 1) * HW + q];
          const float dist = decode_range(x0, p->sigma_mod, geo.recip);
          double P[3];
          unproject(dist, b->cos_az[c], b->sin_az[c], b->cos_el[r], b->sin_el[r], P);
          double wx, wy, wz, ww = 1.0;
          if (p->variant == SDPC_VARIANT_POSE) {
            const double* m = b->to_world + (size_t)bsrc * 16;
            wx = dot4(m, P[0], P[1], P[2], 1.0); wy = dot4(m + 4, P[0], P[1], P[2], 1.0);
            wz = dot4(m + 8, P[0], P[1], P[2], 1.0); ww = dot4(m + 12, P[0], P[1], P[2], 1.0);
          } else {
            wx = P[0] + (double)b->origins[sa * 3]; wy = P[1] + (double)b->origins[sa * 3 + 1];
            wz = P[2] + (double)b->origins[sa * 3 + 2];
          }
          const unsigned src_id = (unsigned)(sa * HW + q);
          for (int t = t_lo; t < t_hi; ++t) {
            double qx, qy, qz;
            if (p->variant == SDPC_VARIANT_POSE) {
              const double* m = b->from_world + (size_t)t * 16;
              qx = dot4(m, wx, wy, wz, ww); qy = dot4(m + 4, wx, wy, wz, ww); qz = dot4(m + 8, wx, wy, wz, ww);
            } else {
              const int ta = t - g * A;
              qx = wx - (double)b->origins[ta * 3]; qy = wy - (double)b->origins[ta * 3 + 1];
              qz = wz - (double)b->origins[ta * 3 + 2];
            }
            Candidate cd = reproject(qx, qy, qz, p->sigma_mod, geo);
            {   // the guarded fp32 estimates of the production kernel must give the identical candidate
              int frow, fcol;
              const bool fok = pixel_fast(qx, qy, qz, geo, &frow, &fcol);
              const double r2 = range2(qx, qy, qz);
              const double fnd = log_range_of_r2(r2, p->sigma_mod, geo);            // what resolve evaluates per cell
              const double tnd = fast_log_range_of_r2(r2, p->sigma_mod, geo, log2_tab);   // what the depth sum takes
              if (fok != in_grid(cd, geo) || (fok && (frow != cd.row || fcol != cd.col)) || memcmp(&fnd, &cd.nd, 8) != 0 ||
                  !(fabs(tnd - cd.nd) <= 8e-15 * (1.0 + cd.nd)))
                ++fast_mismatch;
            }
            bool ok = src_ok && in_grid(cd, geo);
            if (p->min_depth_thr >= 0.0f) ok = ok && cd.nd > (double)p->min_depth_thr;
            if (pass == 0 && b->dbg_row) {
              size_t k = (size_t)t * A * HW + src_id;
              b->dbg_row[k] = cd.row; b->dbg_col[k] = cd.col; b->dbg_valid[k] = ok;
            }
            if (!ok) continue;
            size_t cell = ((size_t)t * R + cd.row) * W + cd.col;
            unsigned long long key;
            memcpy(&key, &cd.nd, 8);
            if (pass == 0) {
              zmin[cell] = std::min(zmin[cell], key);
              r2min[cell] = std::min(r2min[cell], range2(qx, qy, qz));
              cnt[cell]++;
              sum_d[cell] += depth_to_fixed(cd.nd);
              sum_i[cell] += inten_to_fixed(x1);
            } else if (zmin[cell] == key) {
              winner[cell] = std::min(winner[cell], src_id);
            }
          }
        }
      }
  for (size_t k = 0; k < cells; ++k) {
    if (b->dbg_cnt) b->dbg_cnt[k] = (int)cnt[k];
    if (b->dbg_winner) b->dbg_winner[k] = cnt[k] ? (int)winner[k] : -1;
    if (b->dbg_min_d) { double d = 0; if (cnt[k]) memcpy(&d, &zmin[k], 8); b->dbg_min_d[k] = d; }
  }
  // resolve
  std::vector<float> img((size_t)B * 2 * HW, 0.0f);
  std::vector<uint8_t> sm((size_t)B * HW, 0);
  for (int t = t0; t < t0 + tn; ++t)
    for (int q = 0; q < HW; ++q) {
      const int r = q / W, c = q % W;
      const size_t i0 = ((size_t)t * 2) * HW + q, i1 = i0 + HW;
      const bool neg = b->x[i0] < 0.0f;
      int gr = neg ? (H - 1 - r) : (r + R - H);
      int gc = neg ? ((c - W / 2 + W) % W) : c;
      size_t cell = ((size_t)t * R + gr) * W + gc;
      double min_d = 0; float min_i = 0;
      if (cnt[cell]) {
        memcpy(&min_d, &zmin[cell], 8);
        unsigned w = winner[cell];
        int wa = w / HW, wp = w % HW;
        min_i = b->x[((size_t)((t / A) * A + wa) * 2 + 1) * HW + wp];
      }
      Fused f = fuse_cell(cnt[cell], sum_d[cell], sum_i[cell], min_d, min_i, p->sigma_mod, p->allowance, geo.recip);
      {   // the production fusion (guarded float32 far test, no pow/log2 round trip) must take the same decision and
          // land within one float32 ulp of the reference arithmetic
        FusedFast ff = fuse_cell_fast(cnt[cell], sum_d[cell], sum_i[cell], cnt[cell] ? r2min[cell] : 0.0, p->sigma_mod,
                                      p->allowance, geo);
        // the nearest log-range is a function of the smallest squared range, bit for bit
        if (cnt[cell]) { const double md = log_range_of_r2(r2min[cell], p->sigma_mod, geo); if (memcmp(&md, &min_d, 8) != 0) ++fast_mismatch; }
        if (ff.far) fuse_far(&ff, min_d, min_i, p->sigma_mod, p->allowance, geo.recip);
        const float a32 = (float)ff.depth, b32 = (float)f.depth;
        const float tol = fabsf(b32) * 1.2e-7f;
        if (fabsf(a32 - b32) > tol || ff.inten != f.inten) ++fast_mismatch;
      }
      img[i0] = (float)(neg ? f.depth * -1.0 : f.depth);
      img[i1] = f.inten;
      sm[(size_t)t * HW + q] = f.filled && b->exist[q] && b->sky[(size_t)t * HW + q];
    }
  if (fast_mismatch) return 100;
  const bool too_high = too_high_gate(mx, p->sigma_mod, geo.recip);
  if (b->too_high) *b->too_high = too_high;
  for (int t = t0; t < t0 + tn; ++t)
    for (int ch = 0; ch < 2; ++ch)
      for (int q = 0; q < HW; ++q) {
        size_t e = ((size_t)t * 2 + ch) * HW + q;
        float corr = too_high ? 0.0f : (float)(-(int)sm[(size_t)t * HW + q] * (b->mask[e] == 0 ? 1 : 0)) * (b->x[e] - img[e]);
        if (b->new_images) b->new_images[e] = img[e];
        b->x[e] = b->x[e] + p->corr_coef * corr;
      }
  return 0;
}

extern "C" int sdpc_langevin_reproject_step(const sdpc_step_params* p, const sdpc_step_buffers* b, void* workspace,
                                            size_t n, void* stream) {
  if (int e = sdpc_langevin_update(p, b, workspace, n, stream)) return e;
  return p->share ? sdpc_crossview_share(p, b, workspace, n, stream) : 0;
}

extern "C" int emul_langevin_reproject_step(const sdpc_step_params* p, const sdpc_step_buffers* b) {
  float ws[64];
  return sdpc_langevin_reproject_step(p, b, ws, sizeof(ws), nullptr);
}
