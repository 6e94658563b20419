// NCSN_LiDAR_small forward on sm_100a: handle, parameter loading, launch plan.
//
// Host-side restatement of the module graph of LiDARGen/models/ncsnv2.py:420-518 with the blocks
// of models/layers.py (ResidualBlock :401-456, RefineBlock :214-249, RCU :112-134, CRP :62-83,
// MSF :165-184, ConvMeanPool :291-313) and InstanceNorm2dPlus (normalization.py:150-176), as a
// flat list of kernel launches over NHWC buffers carved out of the caller's workspace.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <stdlib.h>

#include <functional>
#include <map>
#include <string>
#include <vector>

#include "common.h"
#include "conv_umma.h"
#include "score_kernels.cuh"
#include "score_types.cuh"

namespace sdpc {

struct ParamSlot {
  std::string name;
  std::vector<int64_t> shape;
  size_t numel = 0;
  float* dev = nullptr;
  bool loaded = false;
};

struct ConvW {
  void* w_tc = nullptr;      // [tap][Cout][Cin] in the operand type (bf16 / tf32-rounded fp32)
  void* w_tc_lo = nullptr;   // bf16x3 arm: residual plane bf16(w - hi)
  float* w_simt = nullptr;   // [tap][Cin][Cout] fp32
  const float* bias = nullptr;
  int Cout = 0, Cin = 0, taps = 0;
};

struct Buf {
  char* ptr = nullptr;
  int N = 0, H = 0, W = 0, C = 0, pad = 0, elem = 4;
  size_t lo_off = 0;         // bf16x3 operands: element offset of the residual plane (0 = single plane)
  size_t bytes = 0;
  size_t off = 0;
  bool valid() const { return bytes != 0; }
};

struct Plan {
  void* ws = nullptr;
  int n_views = 0;
  size_t bytes = 0;
  std::vector<std::function<int(cudaStream_t, const float*, const int64_t*, float*)>> ops;
  std::map<std::string, Buf> taps;
  size_t stats_off = 0, stats_bytes = 0;      // zeroed at the start of every forward: stats_kernel sums, tickets
  size_t parts_off = 0;                       // per-tile partial statistics (every slot written before it is read)
  int n_kernels = 0;          // kernels (and memsets) one forward launches
  double umma_flops = 0.0;    // 2*M*N*K summed over the tensor-core convolution launches
  // fixed staging buffers inside the workspace so that the launch sequence can be replayed as a CUDA graph
  float* x_in = nullptr;
  int64_t* labels_in = nullptr;
  float* out_buf = nullptr;
  size_t io_bytes = 0;
  int eager_runs = 0;
  cudaGraphExec_t graph_exec = nullptr;
  void reset_graph() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    graph_exec = nullptr;
    eager_runs = 0;
  }
};

}  // namespace sdpc

using namespace sdpc;

struct sdpc_score {
  sdpc_score_config cfg;
  int num_sms = 148;
  bool keep_all = false;
  bool use_graph = true;      // replay the forward as a CUDA graph (cfg.reserved bit 1 or SDPC_NO_GRAPH=1 disables)
  cudaStream_t cap_stream = nullptr;
  std::vector<ParamSlot> params;
  std::map<std::string, int> index;
  std::map<std::string, ConvW> convs;
  bool finalized = false;
  Plan plan;
  int last_launches = 0;
  double flops_per_view = 0.0;
  // optional CUDA-event bracketing of the tensor-core convolution launches (bench.py roofline)
  bool profiling = false;
  std::vector<cudaEvent_t> ev_pool;
  struct ProfRec { cudaEvent_t a, b; double flops; std::string name; };
  std::vector<ProfRec> prof;
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  int elem_bytes() const {
    return (cfg.precision == SDPC_PREC_BF16 || cfg.precision == SDPC_PREC_BF16X3 || cfg.precision == SDPC_PREC_FP16) ? 2 : 4;
  }
  bool x3() const { return cfg.precision == SDPC_PREC_BF16X3; }
  // single-pass 16-bit arms (bf16 and fp16 operands): packed max-pooling, 16-bit copies of the tensors only a norm pass reads
  bool plain16() const { return cfg.precision == SDPC_PREC_BF16 || cfg.precision == SDPC_PREC_FP16; }
  // the arm's alternate operand format (EpiParams::op_tf32): tf32 rounding of 32-bit operands, half instead of bf16 in
  // 16-bit containers
  int alt_fmt() const { return (cfg.precision == SDPC_PREC_TF32 || cfg.precision == SDPC_PREC_FP16) ? 1 : 0; }
  const float* P(const std::string& n) const { return params[index.at(n)].dev; }
};

namespace sdpc {

// ------------------------------------------------------------------------------------------
// parameter inventory (reference registration order, ncsnv2.py:420-477)
// ------------------------------------------------------------------------------------------
static void add(std::vector<ParamSlot>& v, const std::string& name, std::vector<int64_t> shape) {
  ParamSlot s;
  s.name = name;
  s.shape = shape;
  s.numel = 1;
  for (auto d : shape) s.numel *= (size_t)d;
  v.push_back(s);
}
static void add_norm(std::vector<ParamSlot>& v, const std::string& pre, int c) {
  add(v, pre + ".alpha", {c});
  add(v, pre + ".gamma", {c});
  add(v, pre + ".beta", {c});
}
static void add_conv(std::vector<ParamSlot>& v, const std::string& pre, int co, int ci, int k, bool bias) {
  add(v, pre + ".weight", {co, ci, k, k});
  if (bias) add(v, pre + ".bias", {co});
}
enum ResKind { RES_PLAIN, RES_DOWN_POOL, RES_DILATED };
static void add_res(std::vector<ParamSlot>& v, const std::string& pre, int ci, int co, ResKind kind) {
  if (kind == RES_PLAIN) {
    add_conv(v, pre + ".conv1", co, ci, 3, true);
    add_norm(v, pre + ".normalize2", co);
    add_conv(v, pre + ".conv2", co, co, 3, true);
  } else if (kind == RES_DOWN_POOL) {
    add_conv(v, pre + ".conv1", ci, ci, 3, true);
    add_norm(v, pre + ".normalize2", ci);
    add_conv(v, pre + ".conv2.conv", co, ci, 3, true);
    add_conv(v, pre + ".shortcut.conv", co, ci, 1, true);
  } else {
    add_conv(v, pre + ".conv1", ci, ci, 3, true);
    add_norm(v, pre + ".normalize2", ci);
    add_conv(v, pre + ".conv2", co, ci, 3, true);
    add_conv(v, pre + ".shortcut", co, ci, 3, true);
  }
  add_norm(v, pre + ".normalize1", ci);
}
static void add_refine(std::vector<ParamSlot>& v, const std::string& pre, std::vector<int> in_planes, int f, bool start,
                       bool end) {
  for (size_t i = 0; i < in_planes.size(); ++i)
    for (int b = 1; b <= 2; ++b)
      for (int s = 1; s <= 2; ++s)
        add_conv(v, pre + ".adapt_convs." + std::to_string(i) + "." + std::to_string(b) + "_" + std::to_string(s) + "_conv",
                 in_planes[i], in_planes[i], 3, false);
  for (int b = 1; b <= (end ? 3 : 1); ++b)
    for (int s = 1; s <= 2; ++s)
      add_conv(v, pre + ".output_convs." + std::to_string(b) + "_" + std::to_string(s) + "_conv", f, f, 3, false);
  if (!start)
    for (size_t i = 0; i < in_planes.size(); ++i) add_conv(v, pre + ".msf.convs." + std::to_string(i), f, in_planes[i], 3, true);
  for (int i = 0; i < 2; ++i) add_conv(v, pre + ".crp.convs." + std::to_string(i), f, f, 3, false);
}

static std::vector<ParamSlot> inventory(const sdpc_score_config& c) {
  std::vector<ParamSlot> v;
  const int g = c.ngf, g2 = 2 * c.ngf;
  add(v, "sigmas", {c.num_classes});
  add_conv(v, "begin_conv", g, c.channels + 2, 3, true);
  add_norm(v, "normalizer", g);
  add_conv(v, "end_conv", c.channels, g, 3, true);
  add_res(v, "res1.0", g, g, RES_PLAIN);
  add_res(v, "res1.1", g, g, RES_PLAIN);
  add_res(v, "res2.0", g, g2, RES_DOWN_POOL);
  add_res(v, "res2.1", g2, g2, RES_PLAIN);
  add_res(v, "res3.0", g2, g2, RES_DILATED);
  add_res(v, "res3.1", g2, g2, RES_PLAIN);
  add_res(v, "res4.0", g2, g2, RES_DILATED);
  add_res(v, "res4.1", g2, g2, RES_PLAIN);
  add_refine(v, "refine1", {g2}, g2, true, false);
  add_refine(v, "refine2", {g2, g2}, g2, false, false);
  add_refine(v, "refine3", {g2, g2}, g, false, false);
  add_refine(v, "refine4", {g, g}, g, false, true);
  return v;
}

// ------------------------------------------------------------------------------------------
// plan builder
// ------------------------------------------------------------------------------------------
struct Builder {
  sdpc_score* h;
  Plan* plan;
  char* base;        // nullptr in the sizing pass
  int N;
  size_t top = 0, high = 0;
  std::multimap<size_t, size_t> free_list;   // bytes -> offset
  int status = SDPC_OK;
  double flops = 0.0;

  bool dry() const { return base == nullptr; }
  static size_t al(size_t v) { return (v + 1023) / 1024 * 1024; }

  Buf alloc(int n, int H, int W, int C, int pad, int elem) {
    Buf b;
    b.N = n; b.H = H; b.W = W; b.C = C; b.pad = pad; b.elem = elem;
    b.bytes = al((size_t)n * (H + 2 * pad) * (W + 2 * pad) * C * elem);
    auto it = free_list.find(b.bytes);
    if (it != free_list.end()) {
      b.off = it->second;
      free_list.erase(it);
    } else {
      b.off = top;
      top += b.bytes;
      if (top > high) high = top;
    }
    b.ptr = base ? base + b.off : nullptr;
    return b;
  }
  Buf raw(int H, int W, int C) { return alloc(N, H, W, C, 0, 4); }
  // A tensor that only an InstanceNorm++ -> ELU -> operand pass reads (conv1 of a residual block): the plain bf16 arm
  // stores it in bf16 (its statistics come from the fp32 accumulators in the epilogue), halving that write and re-read.
  Buf raw_h(int H, int W, int C) { return (h->plain16() && !getenv("SDPC_RAW_FP32")) ? alloc(N, H, W, C, 0, 2) : raw(H, W, C); }
  Buf operand(int H, int W, int C, int pad) {
    if (!h->x3()) return alloc(N, H, W, C, pad, h->elem_bytes());
    Buf b = alloc(2 * N, H, W, C, pad, 2);                 // hi planes of all views, then lo planes
    b.N = N;
    b.lo_off = (size_t)N * (H + 2 * pad) * (W + 2 * pad) * C;
    return b;
  }
  void release(Buf& b) {
    if (!b.valid()) return;
    if (!h->keep_all) free_list.insert({b.bytes, b.off});
    b.bytes = 0;
  }
  void tap(const std::string& name, const Buf& b) { plan->taps[name] = b; }

  typedef std::function<int(cudaStream_t, const float*, const int64_t*, float*)> Op;
  void push(Op op, int n_kernels = 1) {
    if (dry()) return;
    plan->ops.push_back(std::move(op));
    plan->n_kernels += n_kernels;
  }
  static unsigned blocks(size_t total, int per = 256) { return (unsigned)((total + per - 1) / per); }

  // ---- norm: statistics + coefficients -------------------------------------------------
  size_t stats_cursor = 0, parts_cursor = 0;
  struct NormRef { size_t stats_off, coef_off; };
  // statistics slot: [N][C][2] doubles filled by stats_kernel (parts_per_view == 0), or per-tile partials
  // [N][parts_per_view][C][2] floats written by a convolution epilogue (fused statistics)
  struct StatsRef { size_t off = (size_t)-1; int parts_per_view = 0; bool valid() const { return off != (size_t)-1; } };
  StatsRef new_stats(int C) {
    StatsRef r;
    r.off = stats_cursor;
    stats_cursor += al((size_t)N * C * 2 * sizeof(double));
    return r;
  }
  StatsRef new_part_stats(int C, int parts_per_view) {
    StatsRef r;
    r.off = parts_cursor;
    r.parts_per_view = parts_per_view;
    parts_cursor += al((size_t)N * parts_per_view * C * 2 * sizeof(float));
    return r;
  }
  bool can_fuse_stats() const { return h->cfg.precision != SDPC_PREC_FP32; }
  NormRef norm(const Buf& x, const std::string& pre, const StatsRef* have = nullptr) {
    NormRef r;
    const bool fused = have && have->valid();
    r.stats_off = new_stats(x.C).off;                     // [N][C][2] doubles (filled by stats_kernel or the reducer)
    r.coef_off = stats_cursor;
    stats_cursor += al((size_t)N * x.C * 3 * sizeof(float));
    const size_t ticket_off = stats_cursor;                 // per-image ticket counters of the fused reducer (zeroed with the stats area)
    if (fused) stats_cursor += al((size_t)N * sizeof(unsigned int));
    if (dry()) return r;
    if (!fused && x.elem != 4) { status = set_error(SDPC_ERR_STATE, "norm %s: a bf16 tensor needs statistics from its producer", pre.c_str()); return r; }
    const float* in = (const float*)x.ptr;
    const int HW = x.H * x.W, C = x.C;
    int chunks = HW / 512;
    if (chunks < 1) chunks = 1;
    if (chunks > 128) chunks = 128;
    const int ppb = HW / chunks;
    const int groups = 256 / (C / 4);
    const size_t smem = (size_t)groups * C * 2 * sizeof(double);
    char* sbase = base + plan->stats_off;
    double* stats = (double*)(sbase + r.stats_off);
    const float* parts = fused ? (const float*)(base + plan->parts_off + have->off) : nullptr;
    const int ppv = fused ? have->parts_per_view : 0;
    unsigned int* tickets = (unsigned int*)(sbase + ticket_off);
    float* coef = (float*)(sbase + r.coef_off);
    const float *al_ = h->P(pre + ".alpha"), *ga = h->P(pre + ".gamma"), *be = h->P(pre + ".beta");
    const int n = N;
    push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
      if (fused) {
        SDPC_CUDA(launch_k(stats_reduce_finalize_kernel, dim3(dim3(C / 8, n)), dim3(256), 0, s, parts, stats, ppv, C, al_, ga, be, coef, HW, tickets));
      } else {
        SDPC_CUDA(launch_k(stats_kernel, dim3(dim3(chunks, n)), dim3(256), smem, s, in, stats, HW, C, ppb));
        SDPC_CUDA(cudaGetLastError());
        SDPC_CUDA(launch_k(norm_finalize_kernel, dim3(n), dim3(C), 0, s, stats, al_, ga, be, coef, HW, C));
      }
      SDPC_CUDA(cudaGetLastError());
      return SDPC_OK;
    }, fused ? 1 : 2);
    return r;
  }
  const float* coef_ptr(const NormRef& r) const { return dry() ? nullptr : (const float*)(base + plan->stats_off + r.coef_off); }

  // ---- operand materialisation ----------------------------------------------------------
  template <typename T>
  void to_operand_t(const Buf& x, const float* coef, Buf& out, int mode, int halo, int tf32) {
    const float* in = (const float*)x.ptr;
    T* o = (T*)out.ptr;
    const int n = N, H = x.H, W = x.W, C = x.C, P = out.pad;
    const size_t total = (size_t)n * H * (W / (x.elem == 2 ? op_pix<__nv_bfloat16>() : op_pix<float>())) * (C / 8);
    const size_t lo_off = out.lo_off;
    const int nz = lo_off ? 2 * n : n;                      // both planes get the zero border
    const size_t border = (size_t)nz * ((size_t)(H + 2 * P) * (W + 2 * P) - (size_t)H * W) * (C / 8);
    const bool zero = halo == HALO_ZERO && P > 0;
    const bool in_h = x.elem == 2;                           // bf16 raw input (see raw_h())
    push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
      if (zero) SDPC_CUDA(launch_k(zero_halo_kernel<T>, dim3(blocks(border)), dim3(256), 0, s, o, nz, H, W, C, P));
      if (in_h) SDPC_CUDA(launch_k(to_operand_kernel<T, SDPC_OP_MINB, __nv_bfloat16>, dim3(blocks(total)), dim3(256), 0, s, (const __nv_bfloat16*)in, coef, o, n, H, W, C, P, mode, halo, tf32, lo_off));
      else SDPC_CUDA(launch_k(to_operand_kernel<T, SDPC_OP_MINB>, dim3(blocks(total)), dim3(256), 0, s, in, coef, o, n, H, W, C, P, mode, halo, tf32, lo_off));
      SDPC_CUDA(cudaGetLastError());
      return SDPC_OK;
    }, zero ? 2 : 1);
  }
  Buf to_operand(const Buf& x, const float* coef, int pad, int mode, int halo, bool force_fp32 = false) {
    Buf out = force_fp32 ? alloc(N, x.H, x.W, x.C, pad, 4) : operand(x.H, x.W, x.C, pad);
    if (dry()) return out;
    const int tf32 = force_fp32 ? 0 : h->alt_fmt();
    if (out.elem == 2) to_operand_t<__nv_bfloat16>(x, coef, out, mode, halo, tf32);
    else to_operand_t<float>(x, coef, out, mode, halo, tf32);
    return out;
  }

  // the plain bf16 arm pools on packed bf16 pairs and lets the CRP convolution leave its out_acc in bf16
  bool pool_h2() const { return h->plain16(); }
  Buf maxpool(const Buf& x, bool elu_in, Buf* x0_out) {
    Buf out = operand(x.H, x.W, x.C, 1);
    if (dry()) return out;
    const float* in = (const float*)x.ptr;
    float* x0 = x0_out ? (float*)x0_out->ptr : nullptr;
    const int n = N, H = x.H, W = x.W, C = x.C;
    const int tf32 = h->alt_fmt();
    void* o = out.ptr;
    const int elem = out.elem, ei = elu_in ? 1 : 0, in_elem = x.elem;
    const size_t lo_off = out.lo_off;
    if (pool_h2()) {
      const unsigned nblk = (unsigned)((size_t)n * (H / kPoolTH) * (W / kPoolTW) * (C / kPoolHCB));
      push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
        if (in_elem == 2) SDPC_CUDA(launch_k(maxpool5_h2_kernel<__nv_bfloat16>, dim3(nblk), dim3(kPoolHThreads), 0, s, (const __nv_bfloat16*)in, x0, (__nv_bfloat16*)o, n, H, W, C, 1, ei, tf32));
        else SDPC_CUDA(launch_k(maxpool5_h2_kernel<float>, dim3(nblk), dim3(kPoolHThreads), 0, s, in, x0, (__nv_bfloat16*)o, n, H, W, C, 1, ei, tf32));
        SDPC_CUDA(cudaGetLastError());
        return SDPC_OK;
      });
      return out;
    }
    if (in_elem != 4) { status = set_error(SDPC_ERR_STATE, "maxpool: fp32 input expected"); return out; }
    const unsigned nblk = (unsigned)((size_t)n * (H / kPoolTH) * (W / kPoolTW) * (C / kPoolCB));
    push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
      if (elem == 2) SDPC_CUDA(launch_k(maxpool5_kernel<__nv_bfloat16>, dim3(nblk), dim3(kPoolThreads), 0, s, in, x0, (__nv_bfloat16*)o, n, H, W, C, 1, ei, tf32, lo_off));
      else SDPC_CUDA(launch_k(maxpool5_kernel<float>, dim3(nblk), dim3(kPoolThreads), 0, s, in, x0, (float*)o, n, H, W, C, 1, ei, tf32, lo_off));
      SDPC_CUDA(cudaGetLastError());
      return SDPC_OK;
    });
    return out;
  }

  // ---- convolution ----------------------------------------------------------------------
  // in: operand with halo (>= dil); epilogue pointers taken from the given buffers.
  void conv(const Buf& in, const std::string& wname, int dil, bool use_bias, const Buf* residual, Buf* out_raw,
            Buf* out_acc, Buf* out_op, bool op_elu, StatsRef* stats_out = nullptr) {
    if (stats_out) {
      *stats_out = StatsRef();
      if (can_fuse_stats() && out_raw) {
        const int co = h->convs.at(wname).Cout;
        const int tile_px = conv_umma_tile_pixels(co);
        const int parts = conv_umma_stats_parts(co);
        *stats_out = new_part_stats(co, (in.H * in.W / tile_px) * parts);
      }
    }
    const ConvW& cw = h->convs.at(wname);
    flops += 2.0 * (double)in.H * in.W * cw.Cout * cw.Cin * cw.taps;
    if (dry()) return;
    ConvGeom g;
    g.N = N; g.H = in.H; g.W = in.W; g.Cin = cw.Cin; g.Cout = cw.Cout; g.taps = cw.taps; g.dil = dil; g.in_pad = in.pad;
    const int tile_px = (h->cfg.precision == SDPC_PREC_FP32) ? 128 : conv_umma_tile_pixels(cw.Cout);
    g.BW = in.W < tile_px ? in.W : tile_px;
    g.BH = tile_px / g.BW;
    g.bw_shift = 0;
    while ((1 << g.bw_shift) < g.BW) ++g.bw_shift;
    g.tiles_w = in.W / g.BW;
    g.tiles_h = in.H / g.BH;
    g.num_tiles = N * g.tiles_w * g.tiles_h;
    if (h->cfg.precision != SDPC_PREC_FP32 && tile_px == 256) g.num_tiles *= cw.Cout / 128;   // swapped: one tile per 128-channel half
    EpiParams e;
    e.bias = use_bias ? cw.bias : nullptr;
    e.residual = residual ? (const float*)residual->ptr : nullptr;
    e.out_raw = out_raw ? (float*)out_raw->ptr : nullptr;
    e.raw_bf16 = (out_raw && out_raw->elem == 2) ? 1 : 0;
    e.out_acc = out_acc ? (void*)out_acc->ptr : nullptr;
    e.acc_bf16 = (out_acc && out_acc->elem == 2) ? 1 : 0;
    e.out_op = out_op ? out_op->ptr : nullptr;
    e.op_pad = out_op ? out_op->pad : 0;
    e.op_elu = op_elu ? 1 : 0;
    e.op_tf32 = h->alt_fmt();
    e.op_lo_off = out_op ? out_op->lo_off : 0;
    e.prefetch_residual = getenv("SDPC_NO_PREFETCH") ? 0 : 1;
#ifdef SDPC_DEV_HOOKS
    e.dev_wrap = getenv("SDPC_DEV_WRAP") ? atoi(getenv("SDPC_DEV_WRAP")) : 0;
    // development timing probe (tools/gpu_epi_probe.sh; build with SDPC_DEV_HOOKS=1): drops parts of the epilogue,
    // results are garbage.  Not compiled into the shipped library.
    if (const char* dd = getenv("SDPC_DEV_EPI_DROP")) {
      const int m = atoi(dd);
      if (m & 1) e.residual = nullptr;
      if (m & 2) { e.out_raw = nullptr; e.out_acc = nullptr; }
      if (m & 4) e.out_op = nullptr;
      if (m & 16) e.op_elu = 0;
      if (m & 32) e.bias = nullptr;
    }
#endif
    g.passes = h->x3() ? 3 : 1;
    e.stats = (stats_out && stats_out->valid()) ? (float*)(base + plan->parts_off + stats_out->off) : nullptr;
#ifdef SDPC_DEV_HOOKS
    if (const char* dd = getenv("SDPC_DEV_EPI_DROP")) { if (atoi(dd) & 8) e.stats = nullptr; }
#endif
    if (cw.taps == 9 && in.pad < dil) { status = set_error(SDPC_ERR_STATE, "plan: halo %d < dilation %d for %s", in.pad, dil, wname.c_str()); return; }
    if (h->cfg.precision == SDPC_PREC_FP32) {
      const float* inp = (const float*)in.ptr;
      const float* w = cw.w_simt;
      const int bw = g.W < 64 ? g.W : 64, bh = 64 / bw;
      if (g.W % bw || g.H % bh || g.Cout % 64 || g.Cin % 16) { status = set_error(SDPC_ERR_UNSUPPORTED, "conv_simt: shape"); return; }
      dim3 grid(N * (g.W / bw) * (g.H / bh), g.Cout / 64);
      push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
        SDPC_CUDA(launch_k(conv_simt_kernel, dim3(grid), dim3(256), 0, s, inp, w, g, e));
        SDPC_CUDA(cudaGetLastError());
        return SDPC_OK;
      });
      return;
    }
    if (g.H % g.BH || g.W % g.BW) { status = set_error(SDPC_ERR_UNSUPPORTED, "conv_umma: H=%d W=%d not tileable", g.H, g.W); return; }
    UmmaConvLaunch L;
    L.geom = g; L.epi = e; L.elem_bytes = h->elem_bytes(); L.num_sms = h->num_sms;
    const int bk = 128 / L.elem_bytes;
    uint64_t adims[4] = {(uint64_t)cw.Cin, (uint64_t)(in.W + 2 * in.pad), (uint64_t)(in.H + 2 * in.pad), (uint64_t)N};
    uint32_t abox[4] = {(uint32_t)bk, (uint32_t)g.BW, (uint32_t)g.BH, 1u};
    uint64_t bdims[3] = {(uint64_t)cw.Cin, (uint64_t)cw.Cout, (uint64_t)cw.taps};
    uint32_t bbox[3] = {(uint32_t)bk, (uint32_t)conv_umma_weight_rows(cw.Cout), 1u};
    if (int st = make_tmap(&L.tmap_a, in.ptr, L.elem_bytes, 4, adims, abox)) { status = st; return; }
    if (int st = make_tmap(&L.tmap_b, cw.w_tc, L.elem_bytes, 3, bdims, bbox)) { status = st; return; }
    L.tmap_a_lo = L.tmap_a;
    L.tmap_b_lo = L.tmap_b;
    L.tmap_half = L.tmap_a;
    L.tmap_half_lo = L.tmap_a;
    L.use_cluster = 0;
    if (conv_umma_cluster() && (!h->x3() || conv_umma_x3_cluster()) && tile_px == 256 && g.num_tiles % 2 == 0) {
      // the operand two paired tiles share, with a half-sized box: the activation tile (two channel halves of a
      // Cout = 256 layer) or the weights (two neighbouring pixel tiles of a Cout = 128 layer)
      int st;
      if (cw.Cout == 256) {
        uint32_t hbox[4] = {(uint32_t)bk, (uint32_t)(g.BH >= 2 ? g.BW : g.BW / 2), (uint32_t)(g.BH >= 2 ? g.BH / 2 : 1), 1u};
        st = make_tmap(&L.tmap_half, in.ptr, L.elem_bytes, 4, adims, hbox);
        if (!st && h->x3()) st = make_tmap(&L.tmap_half_lo, in.ptr + in.lo_off * 2, 2, 4, adims, hbox);
      } else {
        uint32_t hbox[3] = {(uint32_t)bk, 64u, 1u};
        st = make_tmap(&L.tmap_half, cw.w_tc, L.elem_bytes, 3, bdims, hbox);
        if (!st && h->x3()) st = make_tmap(&L.tmap_half_lo, cw.w_tc_lo, 2, 3, bdims, hbox);
      }
      if (st) { status = st; return; }
      if (!h->x3()) L.tmap_half_lo = L.tmap_half;
      L.use_cluster = 1;
    }
    if (h->x3()) {
      if (!in.lo_off || !cw.w_tc_lo) { status = set_error(SDPC_ERR_STATE, "bf16x3: operand %s has no residual plane", wname.c_str()); return; }
      if (int st = make_tmap(&L.tmap_a_lo, in.ptr + in.lo_off * 2, 2, 4, adims, abox)) { status = st; return; }
      if (int st = make_tmap(&L.tmap_b_lo, cw.w_tc_lo, 2, 3, bdims, bbox)) { status = st; return; }
    }
    const double fl = 2.0 * (double)N * in.H * in.W * cw.Cout * cw.Cin * cw.taps;
    plan->umma_flops += fl;
    sdpc_score* hh = h;
    char pbuf[160];
    snprintf(pbuf, sizeof pbuf, "%s %dx%d %d->%d t%d d%d%s%s%s%s%s%s", wname.c_str(), in.H, in.W, cw.Cin, cw.Cout, cw.taps, dil,
             e.bias ? " bias" : "", e.residual ? " res" : "", e.out_raw ? " raw" : "", e.out_acc ? " acc" : "",
             e.out_op ? (e.op_elu ? " op+elu" : " op") : "", e.stats ? " stats" : "");
    const std::string pname = pbuf;
    push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
      if (!hh->profiling) return conv_umma_launch(L, s);
      sdpc_score::ProfRec r{hh->get_event(), hh->get_event(), fl, pname};
      cudaEventRecord(r.a, s);
      int st = conv_umma_launch(L, s);
      cudaEventRecord(r.b, s);
      hh->prof.push_back(r);
      return st;
    });
  }

  // ---- blocks ---------------------------------------------------------------------------
  Buf norm_elu_operand(const Buf& x, const std::string& npre, int pad, int halo, bool force_fp32 = false,
                       const StatsRef* have = nullptr) {
    NormRef r = norm(x, npre, have);
    return to_operand(x, coef_ptr(r), pad, OP_NORM_ELU, halo, force_fp32);
  }

  // ResidualBlock (layers.py:401-456); does not release `x`.  If `elu_op` is given (plain / dilated blocks),
  // the last convolution's epilogue also emits ELU(out) as an operand with halo 1 for the RCU that reads it.
  //   x_stats   : statistics of `x` already accumulated by its producer (or null -> stats_kernel)
  //   out_stats : receives the slot the last convolution fills with the statistics of the block output
  Buf residual_block(const std::string& pre, const Buf& x, ResKind kind, int dil, Buf* elu_op = nullptr,
                     const StatsRef* x_stats = nullptr, StatsRef* out_stats = nullptr) {
    const int d = dil ? dil : 1;
    Buf sc;                                             // shortcut branch (raw)
    if (kind == RES_DILATED) {
      Buf xs = to_operand(x, nullptr, d, OP_COPY, HALO_CIRC);
      sc = raw(x.H, x.W, h->convs.at(pre + ".shortcut.weight").Cout);
      conv(xs, pre + ".shortcut.weight", d, true, nullptr, &sc, nullptr, nullptr, false);
      release(xs);
    }
    Buf a1 = norm_elu_operand(x, pre + ".normalize1", d, HALO_CIRC, false, x_stats);
    Buf t1 = raw_h(x.H, x.W, h->convs.at(pre + ".conv1.weight").Cout);
    StatsRef t1_stats;
    conv(a1, pre + ".conv1.weight", d, true, nullptr, &t1, nullptr, nullptr, false, &t1_stats);
    release(a1);
    Buf out;
    if (kind == RES_DOWN_POOL) {
      Buf a2 = norm_elu_operand(t1, pre + ".normalize2", 1, HALO_ZERO, false, &t1_stats);
      release(t1);
      const int co = h->convs.at(pre + ".conv2.conv.weight").Cout;
      Buf t2 = raw(x.H, x.W, co);
      conv(a2, pre + ".conv2.conv.weight", 1, true, nullptr, &t2, nullptr, nullptr, false);
      release(a2);
      // shortcut: meanpool(conv1x1(x)+b) == conv1x1(meanpool(x))+b  (both linear)
      Buf xp = operand(x.H / 2, x.W / 2, x.C, 0);
      pool(x, nullptr, &xp, nullptr);
      Buf s2 = raw(x.H / 2, x.W / 2, co);
      conv(xp, pre + ".shortcut.conv.weight", 1, true, nullptr, &s2, nullptr, nullptr, false);
      release(xp);
      out = raw(x.H / 2, x.W / 2, co);
      pool(t2, &s2, nullptr, &out, out_stats);
      release(t2);
      release(s2);
    } else {
      Buf a2 = norm_elu_operand(t1, pre + ".normalize2", d, HALO_CIRC, false, &t1_stats);
      release(t1);
      out = raw(x.H, x.W, h->convs.at(pre + ".conv2.weight").Cout);
      if (elu_op) *elu_op = operand(out.H, out.W, out.C, 1);
      conv(a2, pre + ".conv2.weight", d, true, kind == RES_DILATED ? &sc : &x, &out, nullptr, elu_op, true, out_stats);
      release(a2);
      if (kind == RES_DILATED) release(sc);
    }
    tap(pre, out);
    return out;
  }

  //   stats_out : with add + out_raw only; receives the slot of out_raw's partial statistics (fused into the pooling pass)
  void pool(const Buf& x, const Buf* add, Buf* out_op, Buf* out_raw, StatsRef* stats_out = nullptr) {
    const int Ho = x.H / 2, Wo = x.W / 2;
    const size_t px_per_block = (size_t)kMpIter * 256 / (x.C / 4);
    const bool fuse = stats_out && can_fuse_stats() && add && out_raw && !out_op && 256 % (x.C / 4) == 0 &&
                      ((size_t)Ho * Wo) % px_per_block == 0;
    if (stats_out) *stats_out = fuse ? new_part_stats(x.C, (int)((size_t)Ho * Wo / px_per_block)) : StatsRef();
    if (dry()) return;
    const float* in = (const float*)x.ptr;
    const float* ad = add ? (const float*)add->ptr : nullptr;
    void* oo = out_op ? out_op->ptr : nullptr;
    float* orr = out_raw ? (float*)out_raw->ptr : nullptr;
    const int elem = out_op ? out_op->elem : 4;
    const size_t lo_off = out_op ? out_op->lo_off : 0;
    const int n = N, H = x.H, W = x.W, C = x.C, tf32 = h->alt_fmt();
    const size_t total = (size_t)n * (H / 2) * (W / 2) * (C / 4);
    if (fuse) {
      float2* sp = (float2*)(base + plan->parts_off + stats_out->off);
      const unsigned nblk = (unsigned)(total / ((size_t)kMpIter * 256));
      push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
        SDPC_CUDA(launch_k(meanpool_add_stats_kernel, dim3(nblk), dim3(256), 0, s, in, ad, orr, sp, n, H, W, C));
        SDPC_CUDA(cudaGetLastError());
        return SDPC_OK;
      });
      return;
    }
    push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
      if (elem == 2) SDPC_CUDA(launch_k(meanpool_kernel<__nv_bfloat16>, dim3(blocks(total)), dim3(256), 0, s, in, ad, (__nv_bfloat16*)oo, orr, n, H, W, C, tf32, lo_off));
      else SDPC_CUDA(launch_k(meanpool_kernel<float>, dim3(blocks(total)), dim3(256), 0, s, in, ad, (float*)oo, orr, n, H, W, C, tf32, lo_off));
      SDPC_CUDA(cudaGetLastError());
      return SDPC_OK;
    });
  }

  // RCUBlock (layers.py:112-134).  `x` is not released.
  //   first_op : ELU(x) operand (halo 1) if a producer already emitted it (consumed here), else built here
  //   final    : RCU_RAW  -> returns the raw output
  //              RCU_RAW_ELU -> returns raw and also *final_op = ELU(out) operand for a following RCU
  //              RCU_OP_COPY -> only *final_op = out as an operand (MSF conv input); returns an invalid Buf
  enum { RCU_RAW = 0, RCU_RAW_ELU = 1, RCU_OP_COPY = 2 };
  Buf rcu(const std::string& pre, const Buf& x, int n_blocks, Buf* first_op = nullptr, int final = RCU_RAW,
          Buf* final_op = nullptr, StatsRef* final_stats = nullptr) {
    Buf cur = x;
    bool cur_owned = false;
    Buf a = first_op ? *first_op : to_operand(x, nullptr, 1, OP_ELU, HALO_CIRC);
    if (first_op) first_op->bytes = 0;                    // ownership moves here
    for (int b = 1; b <= n_blocks; ++b) {
      const std::string p1 = pre + "." + std::to_string(b) + "_1_conv.weight", p2 = pre + "." + std::to_string(b) + "_2_conv.weight";
      Buf a2 = operand(x.H, x.W, x.C, 1);
      conv(a, p1, 1, false, nullptr, nullptr, nullptr, &a2, true);       // ELU fused, feeds conv 2 directly
      release(a);
      const bool more = b < n_blocks;
      const bool want_raw = more || final != RCU_OP_COPY;
      const bool want_op = more || final != RCU_RAW;
      const bool op_elu = more || final == RCU_RAW_ELU;
      Buf nxt, an;
      if (want_raw) nxt = raw(x.H, x.W, x.C);
      if (want_op) an = operand(x.H, x.W, x.C, 1);
      conv(a2, p2, 1, false, &cur, want_raw ? &nxt : nullptr, nullptr, want_op ? &an : nullptr, op_elu,
           more ? nullptr : final_stats);                                                                  // + block input
      release(a2);
      if (cur_owned) release(cur);
      cur = nxt;
      cur_owned = true;
      a = an;
    }
    if (final_op) *final_op = a;
    return cur;
  }

  // CRPBlock (layers.py:62-83); releases `hbuf`.  *elu_op = ELU(out) operand for the RCU that follows.
  Buf crp(const std::string& pre, Buf& hbuf, Buf* elu_op) {
    Buf x0 = raw(hbuf.H, hbuf.W, hbuf.C);
    Buf p0 = maxpool(hbuf, true, &x0);                 // x0 = ELU(h); p0 = maxpool(x0)
    release(hbuf);
    Buf path1 = pool_h2() ? alloc(N, x0.H, x0.W, x0.C, 0, 2) : raw(x0.H, x0.W, x0.C);     // only the next max-pool reads it
    Buf x1 = raw(x0.H, x0.W, x0.C);
    conv(p0, pre + ".convs.0.weight", 1, false, &x0, &x1, &path1, nullptr, false);
    release(p0);
    release(x0);
    Buf p1 = maxpool(path1, false, nullptr);
    release(path1);
    Buf x2 = raw(x1.H, x1.W, x1.C);
    *elu_op = operand(x1.H, x1.W, x1.C, 1);
    conv(p1, pre + ".convs.1.weight", 1, false, &x1, &x2, nullptr, elu_op, true);
    release(p1);
    release(x1);
    return x2;
  }

  // RefineBlock (layers.py:214-249) with MSFBlock (layers.py:165-184); inputs are not released.
  //   in_ops[i] : ELU(xs[i]) operand already emitted by the producer of xs[i] (or null)
  //   out_op    : if non-null, also emit ELU(out) as an operand for the next refine block's adapt RCU
  Buf refine(const std::string& pre, std::vector<const Buf*> xs, std::vector<Buf*> in_ops, int features, int outH,
             int outW, bool start, bool end, Buf* out_op, StatsRef* out_stats = nullptr) {
    Buf hsum;
    if (start) {
      hsum = rcu(pre + ".adapt_convs.0", *xs[0], 2, in_ops[0], RCU_RAW, nullptr);
    } else {
      std::vector<Buf> ms;
      for (size_t i = 0; i < xs.size(); ++i) {
        Buf mi;                                                 // adapt RCU output directly as the MSF conv operand
        rcu(pre + ".adapt_convs." + std::to_string(i), *xs[i], 2, in_ops[i], RCU_OP_COPY, &mi);
        Buf si = raw(mi.H, mi.W, features);
        const bool chain = i > 0 && ms[i - 1].H == si.H;       // same resolution: accumulate in the epilogue
        conv(mi, pre + ".msf.convs." + std::to_string(i) + ".weight", 1, true, chain ? &ms[i - 1] : nullptr, &si, nullptr,
             nullptr, false);
        release(mi);
        if (chain) release(ms[i - 1]);
        ms.push_back(si);
      }
      if (ms.back().H == outH) {
        hsum = ms.back();
      } else {                                                   // refine4: second branch is at half resolution
        hsum = raw(outH, outW, features);
        if (!dry()) {
          const float *a = (const float*)ms[0].ptr, *b = (const float*)ms[1].ptr;
          float* o = (float*)hsum.ptr;
          const int n = N, C = features, hh = ms[1].H, ww = ms[1].W;
          const size_t total = (size_t)n * outH * outW * (C / 4);
          push([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
            SDPC_CUDA(launch_k(upsample_add_kernel, dim3(blocks(total)), dim3(256), 0, s, a, b, o, n, outH, outW, C, hh, ww));
            SDPC_CUDA(cudaGetLastError());
            return SDPC_OK;
          });
        }
        release(ms[0]);
        release(ms[1]);
      }
    }
    tap(pre + ".msf", hsum);
    Buf c_op;
    Buf c = crp(pre + ".crp", hsum, &c_op);
    tap(pre + ".crp", c);
    Buf out = rcu(pre + ".output_convs", c, end ? 3 : 1, &c_op, out_op ? RCU_RAW_ELU : RCU_RAW, out_op, out_stats);
    release(c);
    tap(pre, out);
    return out;
  }

  // ---- whole network (ncsnv2.py:484-518) ---------------------------------------------------
  void build() {
    const sdpc_score_config& c = h->cfg;
    const int H = c.height, W = c.width, g = c.ngf;
    {   // persistent I/O staging (never released): x, labels, out
      Buf xin = alloc(N, H, W, c.channels, 0, 4), lab = alloc(N, 1, 1, 2, 0, 4), ob = alloc(N, H, W, c.channels, 0, 4);
      plan->x_in = (float*)xin.ptr;
      plan->labels_in = (int64_t*)lab.ptr;
      plan->out_buf = (float*)ob.ptr;
      plan->io_bytes = (size_t)N * c.channels * H * W * sizeof(float);
    }
    Buf r0 = raw(H, W, g);
    StatsRef st_r0;                                    // statistics of begin_conv's output, one partial per 64-pixel block
    if (can_fuse_stats()) st_r0 = new_part_stats(g, H * (W / 64));
    if (!dry()) {
      float* o = (float*)r0.ptr;
      float2* sp = st_r0.valid() ? (float2*)(base + plan->parts_off + st_r0.off) : nullptr;
      const float *wg = h->P("begin_conv.weight"), *bs = h->P("begin_conv.bias");
      const int n = N;
      push([=](cudaStream_t s, const float* x, const int64_t*, float*) -> int {
        SDPC_CUDA(launch_k(begin_conv_kernel<128>, dim3(n * H * (W / 64)), dim3(128), 0, s, x, wg, bs, o, sp, n, H, W));
        SDPC_CUDA(cudaGetLastError());
        return SDPC_OK;
      });
    }
    flops += 2.0 * H * W * g * (c.channels + 2) * 9;
    tap("begin_conv", r0);
    // l*_op / r*_op: ELU(layer) operands emitted by the producing convolution for the refine blocks' RCUs
    Buf l1_op, l2_op, l3_op, l4_op, r1_op, r2_op, r3_op;
    StatsRef st_a, st_b;                               // statistics of the running trunk tensor, filled by its producer
    Buf t = residual_block("res1.0", r0, RES_PLAIN, 0, nullptr, st_r0.valid() ? &st_r0 : nullptr, &st_a);
    release(r0);
    Buf l1 = residual_block("res1.1", t, RES_PLAIN, 0, &l1_op, &st_a, &st_b);
    release(t);
    StatsRef st_p;                                     // res2.0's output comes from the pooling kernel, which also leaves its statistics
    t = residual_block("res2.0", l1, RES_DOWN_POOL, 0, nullptr, &st_b, &st_p);
    Buf l2 = residual_block("res2.1", t, RES_PLAIN, 0, &l2_op, st_p.valid() ? &st_p : nullptr, &st_a);
    release(t);
    t = residual_block("res3.0", l2, RES_DILATED, 2, nullptr, &st_a, &st_b);
    Buf l3 = residual_block("res3.1", t, RES_PLAIN, 2, &l3_op, &st_b, &st_a);
    release(t);
    t = residual_block("res4.0", l3, RES_DILATED, 4, nullptr, &st_a, &st_b);
    Buf l4 = residual_block("res4.1", t, RES_PLAIN, 4, &l4_op, &st_b, nullptr);
    release(t);
    Buf r1 = refine("refine1", {&l4}, {&l4_op}, 2 * g, l4.H, l4.W, true, false, &r1_op);
    release(l4);
    Buf r2 = refine("refine2", {&l3, &r1}, {&l3_op, &r1_op}, 2 * g, l3.H, l3.W, false, false, &r2_op);
    release(l3);
    release(r1);
    Buf r3 = refine("refine3", {&l2, &r2}, {&l2_op, &r2_op}, g, l2.H, l2.W, false, false, &r3_op);
    release(l2);
    release(r2);
    StatsRef st_r4;
    Buf r4 = refine("refine4", {&l1, &r3}, {&l1_op, &r3_op}, g, l1.H, l1.W, false, true, nullptr, &st_r4);
    release(l1);
    release(r3);
    // tail: normalizer -> ELU -> end_conv -> / sigma in one kernel reading the raw trunk (fp32 in every arm)
    NormRef nr = norm(r4, "normalizer", &st_r4);
    flops += 2.0 * H * W * c.channels * g * 9;
    if (H % kEndRows) { status = set_error(SDPC_ERR_UNSUPPORTED, "end_conv: H=%d not a multiple of %d", H, kEndRows); return; }
    if (!dry()) {
      const float* rw = (const float*)r4.ptr;
      const float* cf = coef_ptr(nr);
      const float *wg = h->P("end_conv.weight"), *bs = h->P("end_conv.bias"), *sg = h->P("sigmas");
      const int n = N, fast = h->cfg.precision != SDPC_PREC_FP32;
      push([=](cudaStream_t s, const float*, const int64_t* labels, float* out) -> int {
        static bool attr_set[kMaxDevices] = {};
        int dev = 0;
        SDPC_CUDA(cudaGetDevice(&dev));
        if (dev < 0 || dev >= kMaxDevices) return set_error(SDPC_ERR_UNSUPPORTED, "end_conv: device ordinal %d >= %d", dev, kMaxDevices);
        if (!attr_set[dev]) {
          SDPC_CUDA(cudaFuncSetAttribute(end_conv_norm_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kEndSmemBytes));
          attr_set[dev] = true;
        }
        const size_t strips = (size_t)n * (H / kEndRows) * (W / 8);
        const int wpb = kEndThreads / 32;
        SDPC_CUDA(launch_k(end_conv_norm_kernel<128>, dim3((unsigned)((strips + wpb - 1) / wpb)), dim3(kEndThreads), kEndSmemBytes, s, rw, cf, wg, bs, sg, labels, out, n, H, W, fast));
        SDPC_CUDA(cudaGetLastError());
        return SDPC_OK;
      });
    }
    release(r4);
  }
};

static int build_plan(sdpc_score* h, int n_views, char* ws, size_t ws_bytes, Plan* plan, size_t* need) {
  // pass 1: sizes (activation arena high-water mark + statistics area)
  Plan scratch;
  Builder dryb{h, &scratch, nullptr, n_views};
  dryb.build();
  const size_t arena = dryb.high, stats = dryb.stats_cursor, parts = dryb.parts_cursor;
  h->flops_per_view = dryb.flops;
  *need = arena + stats + parts + 1024;
  if (!ws) return SDPC_OK;
  if (ws_bytes < *need) return set_error(SDPC_ERR_WORKSPACE, "score workspace too small: %zu < %zu", ws_bytes, *need);
  char* aligned = (char*)(((uintptr_t)ws + 1023) & ~(uintptr_t)1023);
  plan->ops.clear();
  plan->taps.clear();
  plan->reset_graph();
  plan->n_kernels = 1;
  plan->umma_flops = 0.0;
  plan->ws = ws;
  plan->n_views = n_views;
  plan->stats_off = arena;
  plan->stats_bytes = stats;
  plan->parts_off = arena + stats;
  Builder b{h, plan, aligned, n_views};
  // the statistics area is zeroed at the start of every forward
  char* sbase = aligned + arena;
  plan->ops.push_back([=](cudaStream_t s, const float*, const int64_t*, float*) -> int {
    SDPC_CUDA(cudaMemsetAsync(sbase, 0, stats, s));
    return SDPC_OK;
  });
  b.build();
  if (b.status != SDPC_OK) { plan->ops.clear(); plan->ws = nullptr; return b.status; }
  return SDPC_OK;
}

}  // namespace sdpc

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" int sdpc_score_create(const sdpc_score_config* cfg, sdpc_score_t** out) {
  if (!cfg || !out) return set_error(SDPC_ERR_ARG, "score_create: null argument");
  if (cfg->channels != 2) return set_error(SDPC_ERR_UNSUPPORTED, "channels must be 2");
  if (cfg->ngf != 128) return set_error(SDPC_ERR_UNSUPPORTED, "ngf must be 128 (the only LiDAR configuration)");
  if (cfg->precision < 0 || cfg->precision > SDPC_PREC_FP16) return set_error(SDPC_ERR_ARG, "unknown precision %d", cfg->precision);
  if (cfg->height < 16 || cfg->width < 64 || cfg->width % 64 || cfg->height % 16 || (cfg->width & (cfg->width - 1)))
    return set_error(SDPC_ERR_UNSUPPORTED, "need H %% 16 == 0, W a power of two >= 64 (got %dx%d)", cfg->height, cfg->width);
  if (cfg->num_classes <= 0 || cfg->max_views <= 0) return set_error(SDPC_ERR_ARG, "num_classes/max_views must be positive");
  int dev = 0;
  SDPC_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  SDPC_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (cfg->precision != SDPC_PREC_FP32 && prop.major != 10)
    return set_error(SDPC_ERR_UNSUPPORTED, "tcgen05 precisions need an sm_100 device (found sm_%d%d)", prop.major, prop.minor);
  sdpc_score* h = new sdpc_score();
  h->cfg = *cfg;
  h->num_sms = prop.multiProcessorCount;
  h->keep_all = (cfg->reserved & 1) != 0;
  h->use_graph = (cfg->reserved & 2) == 0 && getenv("SDPC_NO_GRAPH") == nullptr;
  h->params = inventory(*cfg);
  for (size_t i = 0; i < h->params.size(); ++i) h->index[h->params[i].name] = (int)i;
  *out = h;
  return SDPC_OK;
}

extern "C" int sdpc_score_destroy(sdpc_score_t* h) {
  if (!h) return SDPC_OK;
  h->plan.reset_graph();
  if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
  for (auto& r : h->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
  for (auto& p : h->params) if (p.dev) cudaFree(p.dev);
  for (auto& kv : h->convs) { if (kv.second.w_tc) cudaFree(kv.second.w_tc); if (kv.second.w_tc_lo) cudaFree(kv.second.w_tc_lo); if (kv.second.w_simt) cudaFree(kv.second.w_simt); }
  delete h;
  return SDPC_OK;
}

extern "C" int sdpc_score_param_count(const sdpc_score_t* h) { return h ? (int)h->params.size() : 0; }

extern "C" int sdpc_score_param_info(const sdpc_score_t* h, int i, const char** name, int64_t shape[4], int* ndim) {
  if (!h || i < 0 || i >= (int)h->params.size()) return set_error(SDPC_ERR_ARG, "param_info: index out of range");
  const ParamSlot& p = h->params[i];
  if (name) *name = p.name.c_str();
  if (ndim) *ndim = (int)p.shape.size();
  if (shape) for (size_t k = 0; k < p.shape.size() && k < 4; ++k) shape[k] = p.shape[k];
  return SDPC_OK;
}

extern "C" int sdpc_score_load_param(sdpc_score_t* h, const char* name, const float* data, const int64_t* shape, int ndim,
                                     int on_device, void* stream) {
  if (!h || !name || !data) return set_error(SDPC_ERR_ARG, "load_param: null argument");
  auto it = h->index.find(name);
  if (it == h->index.end()) return set_error(SDPC_ERR_NAME, "unknown parameter '%s'", name);
  ParamSlot& p = h->params[it->second];
  if (ndim != (int)p.shape.size()) return set_error(SDPC_ERR_NAME, "'%s': expected %d dims, got %d", name, (int)p.shape.size(), ndim);
  for (int k = 0; k < ndim; ++k)
    if (shape[k] != p.shape[k]) return set_error(SDPC_ERR_NAME, "'%s': dim %d is %lld, expected %lld", name, k, (long long)shape[k], (long long)p.shape[k]);
  if (!p.dev) SDPC_CUDA(cudaMalloc(&p.dev, p.numel * sizeof(float)));
  SDPC_CUDA(cudaMemcpyAsync(p.dev, data, p.numel * sizeof(float), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice,
                            (cudaStream_t)stream));
  p.loaded = true;
  h->finalized = false;
  return SDPC_OK;
}

extern "C" int sdpc_score_finalize(sdpc_score_t* h, void* stream_) {
  if (!h) return set_error(SDPC_ERR_ARG, "finalize: null handle");
  cudaStream_t stream = (cudaStream_t)stream_;
  for (auto& p : h->params)
    if (!p.loaded) return set_error(SDPC_ERR_STATE, "parameter '%s' was never loaded", p.name.c_str());
  for (auto& p : h->params) {
    if (p.shape.size() != 4 || p.name == "begin_conv.weight" || p.name == "end_conv.weight") continue;
    ConvW& cw = h->convs[p.name];
    cw.Cout = (int)p.shape[0]; cw.Cin = (int)p.shape[1]; cw.taps = (int)(p.shape[2] * p.shape[3]);
    const std::string bname = p.name.substr(0, p.name.size() - 6) + "bias";
    auto bi = h->index.find(bname);
    cw.bias = bi != h->index.end() ? h->params[bi->second].dev : nullptr;
    const size_t n = p.numel;
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (h->cfg.precision == SDPC_PREC_FP32) {
      if (!cw.w_simt) SDPC_CUDA(cudaMalloc(&cw.w_simt, n * sizeof(float)));
      pack_weight_kernel<float><<<blocks, 256, 0, stream>>>(p.dev, nullptr, cw.w_simt, cw.Cout, cw.Cin, cw.taps, 0);
    } else if (h->cfg.precision == SDPC_PREC_TF32) {
      if (!cw.w_tc) SDPC_CUDA(cudaMalloc(&cw.w_tc, n * sizeof(float)));
      pack_weight_kernel<float><<<blocks, 256, 0, stream>>>(p.dev, (float*)cw.w_tc, nullptr, cw.Cout, cw.Cin, cw.taps, 1);
    } else {
      if (!cw.w_tc) SDPC_CUDA(cudaMalloc(&cw.w_tc, n * sizeof(__nv_bfloat16)));
      if (h->x3() && !cw.w_tc_lo) SDPC_CUDA(cudaMalloc(&cw.w_tc_lo, n * sizeof(__nv_bfloat16)));
      pack_weight_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(p.dev, (__nv_bfloat16*)cw.w_tc, nullptr, cw.Cout, cw.Cin, cw.taps,
                                                                    h->alt_fmt(), (__nv_bfloat16*)cw.w_tc_lo);
    }
    SDPC_CUDA(cudaGetLastError());
  }
  h->finalized = true;
  h->plan.ws = nullptr;
  size_t need = 0;
  Plan tmp;
  return build_plan(h, 1, nullptr, 0, &tmp, &need);      // also fills flops_per_view
}

extern "C" size_t sdpc_score_workspace_bytes(const sdpc_score_t* h_, int n_views) {
  sdpc_score* h = const_cast<sdpc_score*>(h_);
  if (!h || !h->finalized || n_views <= 0) return 0;
  size_t need = 0;
  Plan tmp;
  const double f = h->flops_per_view;
  build_plan(h, n_views, nullptr, 0, &tmp, &need);
  h->flops_per_view = f > 0 ? f : h->flops_per_view;
  return need;
}

extern "C" int sdpc_score_forward(sdpc_score_t* h, const float* x, const int64_t* labels, float* out, int n_views,
                                  void* workspace, size_t workspace_bytes, void* stream_) {
  if (!h || !x || !labels || !out) return set_error(SDPC_ERR_ARG, "forward: null argument");
  if (!h->finalized) return set_error(SDPC_ERR_STATE, "forward before finalize");
  if (n_views <= 0 || n_views > h->cfg.max_views) return set_error(SDPC_ERR_ARG, "n_views %d outside [1, max_views=%d]", n_views, h->cfg.max_views);
  if (!workspace) return set_error(SDPC_ERR_WORKSPACE, "forward: null workspace");
  if (h->plan.ws != workspace || h->plan.n_views != n_views || h->plan.bytes > workspace_bytes) {
    size_t need = 0;
    if (int st = build_plan(h, n_views, (char*)workspace, workspace_bytes, &h->plan, &need)) return st;
    h->plan.bytes = need;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  Plan& pl = h->plan;
  if (!h->use_graph || h->profiling || h->keep_all) {          // eager launches on the caller's buffers
    for (auto& op : pl.ops)
      if (int st = op(stream, x, labels, out)) return st;
  } else {
    // graph path: the kernels only ever see the fixed staging buffers, so one captured graph serves every call
    SDPC_CUDA(cudaMemcpyAsync(pl.x_in, x, pl.io_bytes, cudaMemcpyDeviceToDevice, stream));
    SDPC_CUDA(cudaMemcpyAsync(pl.labels_in, labels, (size_t)n_views * sizeof(int64_t), cudaMemcpyDeviceToDevice, stream));
    if (!pl.graph_exec && pl.eager_runs < 1) {               // first call: eager (sets function attributes, warms up)
      for (auto& op : pl.ops)
        if (int st = op(stream, pl.x_in, pl.labels_in, pl.out_buf)) return st;
      ++pl.eager_runs;
    } else {
      if (!pl.graph_exec) {
        // capture on a private stream: the caller's stream may be the legacy default stream, which cannot capture
        if (!h->cap_stream) SDPC_CUDA(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        cudaGraph_t graph = nullptr;
        SDPC_CUDA(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
        int st = SDPC_OK;
        for (auto& op : pl.ops)
          if ((st = op(h->cap_stream, pl.x_in, pl.labels_in, pl.out_buf))) break;
        cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
        if (st) { if (graph) cudaGraphDestroy(graph); return st; }
        if (ce != cudaSuccess) return set_error(SDPC_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(ce));
        ce = cudaGraphInstantiate(&pl.graph_exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ce != cudaSuccess) { pl.graph_exec = nullptr; return set_error(SDPC_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(ce)); }
      }
      SDPC_CUDA(cudaGraphLaunch(pl.graph_exec, stream));
    }
    SDPC_CUDA(cudaMemcpyAsync(out, pl.out_buf, pl.io_bytes, cudaMemcpyDeviceToDevice, stream));
  }
  h->last_launches = h->plan.n_kernels;
  return SDPC_OK;
}

extern "C" int sdpc_score_read_tap(sdpc_score_t* h, const char* tap, float* out, size_t capacity, int n_views, int chw[3],
                                   void* stream) {
  if (!h || !tap || !out) return set_error(SDPC_ERR_ARG, "read_tap: null argument");
  if (!h->keep_all) return set_error(SDPC_ERR_STATE, "read_tap needs a handle created with reserved|=1 (keep intermediates)");
  auto it = h->plan.taps.find(tap);
  if (it == h->plan.taps.end() || !it->second.ptr) return set_error(SDPC_ERR_NAME, "no tap '%s' (run a forward first)", tap);
  const Buf& b = it->second;
  const size_t total = (size_t)n_views * b.H * b.W * b.C;
  if (capacity < total) return set_error(SDPC_ERR_ARG, "read_tap: capacity %zu < %zu", capacity, total);
  if (chw) { chw[0] = b.C; chw[1] = b.H; chw[2] = b.W; }
  nhwc_to_nchw_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const float*)b.ptr, out, n_views, b.H, b.W, b.C);
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

extern "C" int sdpc_score_last_launch_count(const sdpc_score_t* h) { return h ? h->last_launches : 0; }
extern "C" double sdpc_score_flops_per_view(const sdpc_score_t* h) { return h ? h->flops_per_view : 0.0; }

extern "C" int sdpc_score_set_profiling(sdpc_score_t* h, int on) {
  if (!h) return set_error(SDPC_ERR_ARG, "set_profiling: null handle");
  h->profiling = on != 0;
  return SDPC_OK;
}

extern "C" int sdpc_score_profile_collect(sdpc_score_t* h, double* total_ms, double* total_flops, int* n_launches) {
  if (!h) return set_error(SDPC_ERR_ARG, "profile_collect: null handle");
  double ms = 0.0, fl = 0.0;
  const char* dump = getenv("SDPC_PROFILE_DUMP");             // per-launch csv (name, flops, ms) for tools/conv_layers.py
  FILE* df = dump ? fopen(dump, "a") : nullptr;
  for (auto& r : h->prof) {
    float t = 0.0f;
    cudaError_t ce = cudaEventSynchronize(r.b);
    if (ce == cudaSuccess) ce = cudaEventElapsedTime(&t, r.a, r.b);
    if (ce != cudaSuccess) {
      if (df) fclose(df);
      return set_error(SDPC_ERR_CUDA, "profile_collect: %s", cudaGetErrorString(ce));
    }
    if (df) fprintf(df, "%s,%.0f,%.6f\n", r.name.c_str(), r.flops, t);
    ms += t;
    fl += r.flops;
  }
  if (df) fclose(df);
  for (auto& r : h->prof) {                                   // only now: an error above leaves every record owned by `prof`
    h->ev_pool.push_back(r.a);
    h->ev_pool.push_back(r.b);
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (n_launches) *n_launches = (int)h->prof.size();
  h->prof.clear();
  return SDPC_OK;
}
