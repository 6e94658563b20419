// Implicit-GEMM 3x3 / 1x1 convolution on the 5th-generation tensor cores (tcgen05 + TMEM),
// operands staged by TMA.  This is the kernel that carries ~99% of the score network's FLOPs
// (NCSN_LiDAR_small: LiDARGen/models/ncsnv2.py:484-518, conv definitions layers.py:37-60).
//
// GEMM view:  D[co, m] = sum_{tap, ci} Wt[tap, co, ci] * X[pixel(m) + offset(tap), ci]   (one tcgen05.mma shape: 128 x 256 x 16)
//   M side : 128 output channels (all of a Cout = 128 layer, one half of a Cout = 256 layer)
//   N side : 256 output pixels = a BH x BW box of one image (BW = min(W, 256))
//   K loop : taps x (Cin / BK), BK = 128 bytes of channels (64 bf16 / 32 tf32)
//   (the unswapped tiling - 128 pixels on M, 256 channels on N - is kept behind SDPC_SWAP256=0 for Cout = 256)
// Activations: the input is an NHWC tensor with a materialised halo (circular wrap or zeros), so every shifted
//   tap is an in-bounds 4-D TMA box {BK, BW, BH, 1}; the box lands in shared memory as 256 rows x 128 B,
//   128B-swizzled = the canonical K-major UMMA layout.
// Weights: repacked to [tap][Cout][Cin] (K-major), 3-D TMA box {BK, 128, 1}.
// Clusters: CTAs run in pairs on tiles that share one operand (the activation tile for the two channel halves, the
//   weights for two neighbouring pixel tiles); each CTA fetches half of it and TMA-multicasts it into both.
//   Cout = 256: the pair is driven by ONE tcgen05.mma.cta_group::2 (M = 256 = both channel halves) issued by the leader,
//   the activation tile split between the two CTAs' shared memories instead of multicast (template parameter CG = 2).
// Accumulator: fp32 in TMEM, double buffered (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of
//   tile i+1.  Persistent CTAs, one per SM, static round-robin tiles.
// Warp roles: warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane) + TMEM allocator,
//   warps 2-9 = epilogue (tcgen05.ld -> bias / residual / ELU / statistics / stores, see score_types.cuh).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.h"
#include "conv_umma.h"
#include "score_types.cuh"
#include "sm100_prims.cuh"

namespace sdpc {
using namespace sm100;

constexpr int kTileM = 128;
constexpr int kRowBytes = 128;                  // bytes of K per smem row (one swizzle span)
constexpr int kABytes = kTileM * kRowBytes;     // 16 KiB
constexpr int kThreads = 320;                  // 2 pipeline warps + 8 epilogue warps
constexpr int kSmemBudget = 200 * 1024;
constexpr int kBarrierBytes = 256;
constexpr int kStgFloats = 32 * 32;             // per epilogue warp: 32 pixels x 32 channels, XOR-swizzled float4 slots
constexpr int kEpiWarps = 8;

// CG = 2 (CTA pair driven by one cta_group::2 MMA): each CTA stages only its half of the N-side operand
template <int N_TILE, int CG = 1>
struct UmmaCfg {
  static constexpr int kBBytes = N_TILE / CG * kRowBytes;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // CTA pairs are swapped-operand kernels: their epilogue never touches the transposing staging area, so its 32 KB hold a
  // seventh 32 KB stage instead (7 x 32 KB + 1 KB + 256 B = 225.25 KB, the size the other variants already use)
  static constexpr int kStages = CG == 2 ? 7 : kSmemBudget / kStageBytes;   // 4 (N=256) / 6 (N=128) / 7 (N=256 split over a CTA pair)
  static constexpr int kTmemCols = 2 * N_TILE;                       // 512 / 256
  static constexpr int kStgBytes = CG == 2 ? 0 : kEpiWarps * kStgFloats * 4;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + kBarrierBytes + kStgBytes;
  static_assert(kSmemBytes <= 227 * 1024, "dynamic shared memory of one CTA");
  static_assert((2 * kStages + 4) * 8 + 4 <= kBarrierBytes, "barrier block");
};

// SWAP = false: D[pixel (M=128), cout (N=Cout)]      = X_tile . W^T      (Cout = 256 layers)
// SWAP = true : D[cout  (M=128), pixel (N=256)]      = W . X_tile^T      (Cout = 128 layers; Cout = 256 layers as two
//                                                                          128-channel halves, COUT = 256: same bytes
//                                                                          through L2 per FLOP, cheaper epilogue)
//   A 128x128 MMA reads (128+128) rows of operands per 128x128 MACs and is bound by shared-memory
//   bandwidth; putting the weights on the M side lets a 128-output-channel layer run the same
//   128x256 instruction shape as the 256-channel layers (256 pixels per tile on the N side).
// CL (swapped layers only): the kernel runs as 2-CTA clusters on paired tiles that share one operand - the activation
// tile for the two 128-channel halves of a Cout = 256 layer, the weights for two neighbouring pixel tiles of a
// Cout = 128 layer.  Each CTA fetches half of the shared operand and TMA-multicasts it into both CTAs (tmap_half has the
// half-sized box), so the L2 -> SM operand traffic drops by a third (a sixth); a stage is released to both producers by
// a multicast tcgen05.commit (empty barriers count two arrivals).
// CG = 2 (Cout = 256 clusters; SDPC_CTA2=0 falls back to CG = 1 with multicast): the pair runs ONE tcgen05.mma.cta_group::2 of shape 256 x 256 per k-step
// instead of two 128 x 256 ones.  M = 256 is the two 128-channel halves (one per CTA, each accumulating in its own
// TMEM), the 256-pixel activation tile is split in two halves of 128 rows, one in each CTA's shared memory: nothing is
// multicast, every SM reads 128 + 128 operand rows from its shared memory per k-step instead of 128 + 256, and a stage
// is 32 KB instead of 48 KB (6 stages instead of 4).  Only the leader (cluster rank 0) issues MMAs; both producers
// report their bytes to the leader's full barrier, the leader's commits are multicast to the empty / accumulator-full
// barriers of both CTAs, and the epilogue warps of both CTAs release an accumulator on the leader's barrier.
template <typename T, int N_TILE, bool SWAP, int COUT, bool CL, int CG = 1>
__global__ void __launch_bounds__(kThreads, 1)
conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_a_lo, const __grid_constant__ CUtensorMap tmap_b_lo,
                 const __grid_constant__ CUtensorMap tmap_half, const __grid_constant__ CUtensorMap tmap_half_lo,
                 const ConvGeom g, const EpiParams e) {
  static_assert(!CL || SWAP, "clusters are implemented for the swapped-operand variant");
  static_assert(CG == 1 || (CG == 2 && CL && COUT == 2 * kTileM), "CTA pairs: the two 128-channel halves of a Cout = 256 layer");
  using Cfg = UmmaCfg<N_TILE, CG>;
  constexpr bool kTf32 = sizeof(T) == 4;
  constexpr int kBK = kRowBytes / (int)sizeof(T);     // channels per k-block: 64 bf16 / 32 tf32
  constexpr int kUmmaK = 32 / (int)sizeof(T);         // 16 / 8 -> 32 bytes per MMA along K
  constexpr int kMmasPerStage = kBK / kUmmaK;         // 4
  // operand format of the MMA: tf32 for 32-bit operands; for 16-bit operands bf16 (1) or, in the fp16 arm, half (0)
  const uint32_t kIdesc = umma_idesc(kTf32 ? 2u : (e.op_tf32 ? 0u : 1u), CG * kTileM, N_TILE);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full_bar = bars;                         // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + Cfg::kStages;         // [kStages]  MMA -> TMA
  uint64_t* acc_full = bars + 2 * Cfg::kStages;      // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;                // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    if (g.passes > 1) { prefetch_tmap(&tmap_a_lo); prefetch_tmap(&tmap_b_lo); }
    if (CL) prefetch_tmap(&tmap_half);
    if (CL && g.passes > 1) prefetch_tmap(&tmap_half_lo);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, (CL && CG == 1) ? 2 : 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(acc_full + i, 1); mbar_init(acc_empty + i, CG * kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (CG == 2) { tmem_alloc_cg2(tmem_slot, Cfg::kTmemCols); tmem_relinquish_cg2(); }
    else { tmem_alloc(tmem_slot, Cfg::kTmemCols); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t crank = CL ? cluster_ctarank() : 0;
  if (CL) cluster_sync_all();                          // the peer's barriers are initialised before anything is multicast
  // Everything above (barriers, descriptor prefetch, TMEM allocation) may overlap the previous kernel's tail; its
  // results (this layer's input, residual, the buffers this layer overwrites) are only touched after this point.
  pdl_sync();

  const int k_chunks = g.Cin / kBK;
  const int k_iters = g.passes * g.taps * k_chunks;
  const int tiles_per_img = g.tiles_w * g.tiles_h;
  constexpr int kMH = SWAP ? COUT / kTileM : 1;          // swapped: tiles = pixel tiles x 128-channel halves
  static_assert(SWAP ? (COUT == 128 || COUT == 256) : COUT == N_TILE, "unsupported channel count");

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x) {
        const int ptile = tile / kMH, mh = tile - ptile * kMH;   // pixel tile, 128-channel half of the weights (swapped)
        const int n = ptile / tiles_per_img;
        const int rem = ptile - n * tiles_per_img;
        const int h0 = (rem / g.tiles_w) * g.BH;
        const int w0 = (rem % g.tiles_w) * g.BW;
        // bf16x3 arm: three passes over K accumulate X_hi.W_hi + X_hi.W_lo + X_lo.W_hi into the same accumulator
        for (int pass = 0; pass < g.passes; ++pass) {
          const CUtensorMap* ma = pass == 2 ? &tmap_a_lo : &tmap_a;
          const CUtensorMap* mb = pass == 1 ? &tmap_b_lo : &tmap_b;
          // half box of the operand the pair shares: the activations (lo plane in pass 2) for the two channel halves of
          // a Cout = 256 layer, the weights (lo plane in pass 1) for two pixel tiles of a Cout = 128 layer
          const CUtensorMap* mhalf = (pass == (kMH == 2 ? 2 : 1)) ? &tmap_half_lo : &tmap_half;
          for (int tap = 0; tap < g.taps; ++tap) {
            const int dy = (g.taps == 9) ? (tap / 3 - 1) * g.dil : 0;
            const int dx = (g.taps == 9) ? (tap % 3 - 1) * g.dil : 0;
            for (int kc = 0; kc < k_chunks; ++kc) {
              mbar_wait(empty_bar + stage, phase ^ 1);
              uint8_t* sa = smem + stage * Cfg::kStageBytes;
              uint8_t* sb = sa + kABytes;
              if constexpr (CG == 2) {
                // own weight half and own half of the activation tile; the leader's barrier counts the bytes of both CTAs
                if (crank == 0) mbar_arrive_expect_tx(full_bar + stage, 2 * Cfg::kStageBytes);
                const uint32_t lead_full = mapa_u32(full_bar + stage, 0);
                tma_load_3d_cg2(sa, mb, lead_full, kc * kBK, mh * kTileM, tap);
                const int hw = g.BH >= 2 ? 0 : (int)crank * (g.BW / 2), hh = g.BH >= 2 ? (int)crank * (g.BH / 2) : 0;
                tma_load_4d_cg2(sb, mhalf, lead_full, kc * kBK, w0 + hw + g.in_pad + dx, h0 + hh + g.in_pad + dy, n);
              } else if constexpr (!CL) {
                // activations and weights land in the M-side (128 rows) or N-side (N_TILE rows) slot
                mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
                tma_load_4d(SWAP ? sb : sa, ma, full_bar + stage, kc * kBK, w0 + g.in_pad + dx, h0 + g.in_pad + dy, n);
                tma_load_3d(SWAP ? sa : sb, mb, full_bar + stage, kc * kBK, mh * kTileM, tap);
              } else if constexpr (kMH == 2) {
                // own weight half; rows [128 r, 128 r + 128) of the shared activation tile, multicast to both CTAs
                mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
                tma_load_3d(sa, mb, full_bar + stage, kc * kBK, mh * kTileM, tap);
                const int hw = g.BH >= 2 ? 0 : (int)crank * (g.BW / 2), hh = g.BH >= 2 ? (int)crank * (g.BH / 2) : 0;
                tma_load_4d_mc(sb + crank * (N_TILE / 2) * kRowBytes, mhalf, full_bar + stage, kc * kBK,
                               w0 + hw + g.in_pad + dx, h0 + hh + g.in_pad + dy, n, (uint16_t)3);
              } else {
                // own activation tile; rows [64 r, 64 r + 64) of the shared weights, multicast to both CTAs
                mbar_arrive_expect_tx(full_bar + stage, Cfg::kStageBytes);
                tma_load_4d(sb, ma, full_bar + stage, kc * kBK, w0 + g.in_pad + dx, h0 + g.in_pad + dy, n);
                tma_load_3d_mc(sa + crank * (kTileM / 2) * kRowBytes, mhalf, full_bar + stage, kc * kBK,
                               (int)crank * (kTileM / 2), tap, (uint16_t)3);
              }
              if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    if (lane == 0 && (CG == 1 || crank == 0)) {              // CTA pairs: the leader issues for both SMs
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++local) {
        const int ab = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(acc_empty + ab, acc_phase ^ 1);            // epilogue (of both CTAs) has drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * N_TILE;
        for (int k = 0; k < k_iters; ++k) {
          mbar_wait(full_bar + stage, phase);                // TMA bytes have landed
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t da = umma_desc_sw128_kmajor(sa);
          const uint64_t db = umma_desc_sw128_kmajor(sa + kABytes);
#pragma unroll
          for (int j = 0; j < kMmasPerStage; ++j) {
            // advance 32 bytes along K inside the swizzle span: +2 in the (addr >> 4) field
            if constexpr (CG == 2) umma_ss_cg2<kTf32>(d_tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), kIdesc, (k | j) != 0);
            else umma_ss<kTf32>(d_tmem, da + (uint64_t)(2 * j), db + (uint64_t)(2 * j), kIdesc, (k | j) != 0);
          }
          if constexpr (CG == 2) umma_commit_cg2(empty_bar + stage, (uint16_t)3);  // the slot is free in both CTAs
          else if constexpr (CL) umma_commit_mc(empty_bar + stage, (uint16_t)3);   // both producers write this slot
          else umma_commit(empty_bar + stage);               // frees the smem slot when the MMAs retire
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        if constexpr (CG == 2) umma_commit_cg2(acc_full + ab, (uint16_t)3);   // both halves of the accumulator complete
        else umma_commit(acc_full + ab);                     // accumulator complete
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue (warps 2..9)
    // Eight warps: two per TMEM lane quadrant (a warp may only read lanes 32*(warp%4)..+31), taking the even
    // and the odd 32-column chunks of the accumulator.  TMEM hands a thread one row x 32 columns.
    //  SWAP   : row = output channel, columns = 32 consecutive pixels of one image row.  A warp-wide store of
    //           column j is 32 consecutive channels of one pixel = one full 128-byte line (64 B for bf16), so
    //           the registers are stored as they are: no transpose, no per-element address arithmetic
    //           (pixel stride kCout is a compile-time immediate), per-thread bias and statistics.
    //  !SWAP  : row = pixel, columns = 32 channels; storing that directly would touch 32 different lines per
    //           instruction, so every 32x32 chunk is transposed through an XOR-swizzled shared-memory tile:
    //           8 consecutive lanes then cover one pixel's 128 contiguous bytes.
    constexpr int kCout = COUT;                              // the host only launches this variant for that Cout
    const int quad = warp & 3;                               // TMEM lane quadrant this warp may read
    const int half = (warp - 2) >> 2;                        // 0: even chunks, 1: odd chunks
    float* stg = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + kBarrierBytes) + (warp - 2) * kStgFloats;
    const int cq = lane & 7;                                 // channel quad within the 32-channel chunk
    const int psub = lane >> 3;                              // pixel within a group of 4
    // staging tile: 32 rows (pixels) x 8 float4 (32 channels); float4 slot of (row, c4) = row*8 + (c4 ^ (row & 7))
    auto stg_slot = [](int row, int c4) { return (row << 3) + (c4 ^ (row & 7)); };
    const bool red = e.op_tf32 != 0;
    const int P = e.op_pad, Wp = g.W + 2 * P, Hp = g.H + 2 * P;
    int local = 0;
    for (int tile = blockIdx.x; tile < g.num_tiles; tile += gridDim.x, ++local) {
      const int ab = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int ptile = tile / kMH, mh = tile - ptile * kMH;
      const int n = ptile / tiles_per_img;
      const int rem = ptile - n * tiles_per_img;
      const int h0 = (rem / g.tiles_w) * g.BH;
      const int w0 = (rem % g.tiles_w) * g.BW;
      // While the MMAs of this tile are still running, pull the residual rows of the tile into L2 so that the
      // epilogue's residual loads hit L2 instead of HBM (the 8 warps cover the tile's pixels x Cout floats).
      if (e.residual && e.prefetch_residual) {
        const int tile_px = g.BW * g.BH;
        constexpr int lines_per_px = (SWAP ? kTileM : kCout) / 32;  // 128-byte lines per pixel (of this tile's channels)
        for (int i = (warp - 2) * 32 + lane; i < tile_px * lines_per_px; i += kEpiWarps * 32) {
          const int m = i / lines_per_px, l = i - m * lines_per_px;
          const int h = h0 + (m >> g.bw_shift), w = w0 + (m & (g.BW - 1));
          const float* ptr = e.residual + (((size_t)n * g.H + h) * g.W + w) * kCout + mh * kTileM + l * 32;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        }
      }
      const uint32_t t_addr = tmem_base + ((uint32_t)(quad * 32) << 16) + ab * N_TILE;

      if constexpr (SWAP) {
        const int ch = mh * kTileM + quad * 32 + lane;       // this thread's output channel
        const float bias = e.bias ? e.bias[ch] : 0.0f;
        float ssum = 0.0f, ssq = 0.0f;                       // InstanceNorm++ partial sums of this channel over the warp's pixels
        mbar_wait(acc_full + ab, acc_phase);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int c0 = half * 32; c0 < N_TILE; c0 += 64) {
          // the chunk's 32 pixels lie on one image row (BW is a multiple of 32): pixel j at (h, w + j)
          const int h = h0 + (c0 >> g.bw_shift), w = w0 + (c0 & (g.BW - 1));
          size_t raw0 = (((size_t)n * g.H + h) * g.W + w) * kCout + ch;
          size_t res0 = raw0;
#ifdef SDPC_DEV_HOOKS
          if (e.dev_wrap & 1) raw0 &= (size_t)0x3FFFF;       // timing probes only (results are garbage)
          if (e.dev_wrap & 2) res0 &= (size_t)0x3FFFF;
#endif
          float res[32];
          if (e.residual) {                                  // issued first: the latency overlaps the TMEM read
            const float* rp = e.residual + res0;
#pragma unroll
            for (int j = 0; j < 32; ++j) res[j] = rp[j * kCout];
          }
          uint32_t r[32];
          tmem_ld_32x32(t_addr + c0, r);
          tmem_ld_wait();
          if (c0 + 64 >= N_TILE) {                           // this warp's last chunk: its part of the accumulator is read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(acc_empty + ab, 0));   // the leader's barrier
              else mbar_arrive(acc_empty + ab);
            }
          }
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + bias;
          if (e.out_acc) {
            if (e.acc_bf16) {
              __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(e.out_acc) + raw0;
#pragma unroll
              for (int j = 0; j < 32; ++j) *reinterpret_cast<uint16_t*>(&ap[j * kCout]) = pack1_h16(v[j], e.op_tf32 != 0);
            } else {
              float* ap = reinterpret_cast<float*>(e.out_acc) + raw0;
#pragma unroll
              for (int j = 0; j < 32; ++j) ap[j * kCout] = v[j];
            }
          }
          if (e.residual) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += res[j];
          }
          if (e.out_raw) {
            if (e.raw_bf16) {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out_raw) + raw0;
#pragma unroll
              for (int j = 0; j < 32; ++j) *reinterpret_cast<uint16_t*>(&op[j * kCout]) = pack1_h16(v[j], e.op_tf32 != 0);
            } else {
              float* op = e.out_raw + raw0;
#pragma unroll
              for (int j = 0; j < 32; ++j) op[j * kCout] = v[j];
            }
          }
          if (e.stats) {
#pragma unroll
            for (int j = 0; j < 32; ++j) { ssum += v[j]; ssq = fmaf(v[j], v[j], ssq); }
          }
          if (e.out_op) {
            if (e.op_elu) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = elu_sel<T>(v[j], red);
            }
            size_t op0 = (((size_t)n * Hp + h + P) * Wp + w + P) * kCout + ch;
#ifdef SDPC_DEV_HOOKS
            if (e.dev_wrap & 1) op0 &= (size_t)0x3FFFF;
#endif
            T* const ob = reinterpret_cast<T*>(e.out_op) + op0;
            // pixels [jlo, jhi) of the chunk to `ob + delta`: the chunk itself, then its circular-halo duplicates
            auto emit = [&](ptrdiff_t delta, int jlo, int jhi) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j >= jlo && j < jhi) store_op1<T>(ob + delta + j * kCout, v[j], red, e.op_lo_off);
            };
            emit(0, 0, 32);
            const ptrdiff_t dh = (h < P) ? (ptrdiff_t)g.H * Wp * kCout : ((h >= g.H - P) ? -(ptrdiff_t)g.H * Wp * kCout : 0);
            if (dh) emit(dh, 0, 32);
            if (w < P) {                                     // leftmost pixels also live in the right halo
              emit((ptrdiff_t)g.W * kCout, 0, P - w);
              if (dh) emit(dh + (ptrdiff_t)g.W * kCout, 0, P - w);
            }
            if (w + 32 > g.W - P) {                          // rightmost pixels also live in the left halo
              emit(-(ptrdiff_t)g.W * kCout, g.W - P - w, 32);
              if (dh) emit(dh - (ptrdiff_t)g.W * kCout, g.W - P - w, 32);
            }
          }
        }
        if (e.stats)                                         // slot [tile][half][Cout][2], written exactly once
          *reinterpret_cast<float2*>(e.stats + (((size_t)ptile * 2 + half) * kCout + ch) * 2) = make_float2(ssum, ssq);
      } else {
        // The quadrant's 32 pixels lie on one image row (BW >= 32): the lane's pixel `it` is (h, w + 4*it), so every
        // tensor offset is one per-tile base plus a compile-time multiple of kCout.  Warps whose pixels touch the
        // image border additionally write the circular-halo duplicates of the operand (slow path, warp-uniform).
        const int m0 = quad * 32;
        const int h = h0 + (m0 >> g.bw_shift), wq = w0 + (m0 & (g.BW - 1)), w = wq + psub;
        const size_t raw0 = (((size_t)n * g.H + h) * g.W + w) * kCout;
        const size_t op0 = (((size_t)n * Hp + h + P) * Wp + w + P) * kCout;
        const bool edge = P > 0 && (h < P || h >= g.H - P || wq < P || wq + 32 > g.W - P);
        // fused InstanceNorm++ statistics of the value written to out_raw: per lane 4 channels, summed over the
        // lane's pixels, then over the 4 lanes sharing a channel quad; the warp's partial goes to its own slot
        // [tile][quad][Cout][2] (no atomics, every slot written exactly once; the reducer sums the slots).
        float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
        mbar_wait(acc_full + ab, acc_phase);
        __syncwarp();
        tc_fence_after();
#pragma unroll 1
        for (int c0 = half * 32; c0 < N_TILE; c0 += 64) {
          const int ch = c0 + cq * 4;
          // issue the residual loads first: their latency overlaps the TMEM read and the staging round trip
          float4 res[8];
          if (e.residual) {
            const float* rp = e.residual + raw0 + ch;
#pragma unroll
            for (int it = 0; it < 8; ++it) res[it] = *reinterpret_cast<const float4*>(rp + it * 4 * kCout);
          }
          uint32_t r[32];
          tmem_ld_32x32(t_addr + c0, r);
          tmem_ld_wait();
          if (c0 + 64 >= N_TILE) {                           // this warp's last chunk: its part of the accumulator is read
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if constexpr (CG == 2) mbar_arrive_cluster(mapa_u32(acc_empty + ab, 0));   // the leader's barrier
              else mbar_arrive(acc_empty + ab);
            }
          }
          float4* stg4 = reinterpret_cast<float4*>(stg);
          // thread = pixel (row = lane), registers = 32 channels
#pragma unroll
          for (int j = 0; j < 8; ++j)
            stg4[stg_slot(lane, j)] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                  __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
          __syncwarp();
          float4 v[8];
#pragma unroll
          for (int it = 0; it < 8; ++it) v[it] = stg4[stg_slot(it * 4 + psub, cq)];
          if (e.bias) {
            const float4 b4 = *reinterpret_cast<const float4*>(e.bias + ch);
#pragma unroll
            for (int it = 0; it < 8; ++it) { v[it].x += b4.x; v[it].y += b4.y; v[it].z += b4.z; v[it].w += b4.w; }
          }
          if (e.out_acc) {
            if (e.acc_bf16) {
              __nv_bfloat16* ap = reinterpret_cast<__nv_bfloat16*>(e.out_acc) + raw0 + ch;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const float a4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
                store_op4<__nv_bfloat16>(ap + it * 4 * kCout, a4, e.op_tf32 != 0, 0);
              }
            } else {
              float* ap = reinterpret_cast<float*>(e.out_acc) + raw0 + ch;
#pragma unroll
              for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(ap + it * 4 * kCout) = v[it];
            }
          }
          if (e.residual) {
#pragma unroll
            for (int it = 0; it < 8; ++it) { v[it].x += res[it].x; v[it].y += res[it].y; v[it].z += res[it].z; v[it].w += res[it].w; }
          }
          if (e.out_raw) {
            if (e.raw_bf16) {
              __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(e.out_raw) + raw0 + ch;
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const float a4[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
                store_op4<__nv_bfloat16>(op + it * 4 * kCout, a4, e.op_tf32 != 0, 0);
              }
            } else {
              float* op = e.out_raw + raw0 + ch;
#pragma unroll
              for (int it = 0; it < 8; ++it) *reinterpret_cast<float4*>(op + it * 4 * kCout) = v[it];
            }
          }
          if (e.stats) {
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              ssum[0] += v[it].x; ssum[1] += v[it].y; ssum[2] += v[it].z; ssum[3] += v[it].w;
              ssq[0] = fmaf(v[it].x, v[it].x, ssq[0]); ssq[1] = fmaf(v[it].y, v[it].y, ssq[1]);
              ssq[2] = fmaf(v[it].z, v[it].z, ssq[2]); ssq[3] = fmaf(v[it].w, v[it].w, ssq[3]);
            }
          }
          if (e.out_op) {
            if (e.op_elu) {
#pragma unroll
              for (int it = 0; it < 8; ++it)
                v[it] = make_float4(elu_sel<T>(v[it].x, red), elu_sel<T>(v[it].y, red), elu_sel<T>(v[it].z, red), elu_sel<T>(v[it].w, red));
            }
            T* const ob = reinterpret_cast<T*>(e.out_op) + op0 + ch;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
              const float o[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
              store_op4<T>(ob + it * 4 * kCout, o, red, e.op_lo_off);
            }
            if (edge) {
              const ptrdiff_t dh = (h < P) ? (ptrdiff_t)g.H * Wp * kCout : ((h >= g.H - P) ? -(ptrdiff_t)g.H * Wp * kCout : 0);
#pragma unroll
              for (int it = 0; it < 8; ++it) {
                const int wi = w + it * 4;
                const ptrdiff_t dw = (wi < P) ? (ptrdiff_t)g.W * kCout : ((wi >= g.W - P) ? -(ptrdiff_t)g.W * kCout : 0);
                const float o[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
                T* const d = ob + it * 4 * kCout;
                if (dw) store_op4<T>(d + dw, o, red, e.op_lo_off);
                if (dh) {
                  store_op4<T>(d + dh, o, red, e.op_lo_off);
                  if (dw) store_op4<T>(d + dh + dw, o, red, e.op_lo_off);
                }
              }
            }
          }
          if (e.stats) {                                     // every chunk covers different channels
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              ssum[k] += __shfl_xor_sync(0xffffffffu, ssum[k], 8);
              ssq[k] += __shfl_xor_sync(0xffffffffu, ssq[k], 8);
              ssum[k] += __shfl_xor_sync(0xffffffffu, ssum[k], 16);
              ssq[k] += __shfl_xor_sync(0xffffffffu, ssq[k], 16);
            }
            if (psub == 0) {
              float* sp = e.stats + (((size_t)tile * 4 + quad) * kCout + ch) * 2;
              *reinterpret_cast<float4*>(sp) = make_float4(ssum[0], ssq[0], ssum[1], ssq[1]);
              *reinterpret_cast<float4*>(sp + 4) = make_float4(ssum[2], ssq[2], ssum[3], ssq[3]);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) ssum[k] = ssq[k] = 0.f;
          }
          __syncwarp();                                      // staging tile is reused by the next chunk
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();                     // the pair's TMEM is released together: both epilogues are done
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    if constexpr (CG == 2) tmem_dealloc_cg2(tmem_base, Cfg::kTmemCols);
    else tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
  if (CL) cluster_sync_all();                          // no CTA leaves while its peer may still signal its barriers
}

// ------------------------------------------------------------------------------------------ host
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_tmap(CUtensorMap* out, void* base, int elem_bytes, int rank, const uint64_t* dims, const uint32_t* box) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return set_error(SDPC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t gdim[5], gstride[5];
  cuuint32_t bdim[5], estr[5];
  uint64_t stride = (uint64_t)elem_bytes;
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    stride *= dims[i];
    if (i < rank - 1) gstride[i] = stride;       // byte stride of dimension i+1
  }
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = enc(out, dt, (cuuint32_t)rank, base, gdim, gstride, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(SDPC_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return SDPC_OK;
}

template <typename T, int N_TILE, bool SWAP, int COUT, bool CL, int CG = 1>
static int launch_t(const UmmaConvLaunch& L, cudaStream_t stream) {
  using Cfg = UmmaCfg<N_TILE, CG>;
  auto kernel = conv_umma_kernel<T, N_TILE, SWAP, COUT, CL, CG>;
  static bool attr_set_dev[kMaxDevices] = {};
  static int max_clusters_dev[kMaxDevices] = {};
  int dev = 0;
  SDPC_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= kMaxDevices) return set_error(SDPC_ERR_UNSUPPORTED, "conv_umma: device ordinal %d >= %d", dev, kMaxDevices);
  bool& attr_set = attr_set_dev[dev];
  int& max_clusters = max_clusters_dev[dev];
  if (!attr_set) {
    SDPC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    if (CL) {
      cudaLaunchConfig_t q = {};
      q.gridDim = dim3(L.num_sms & ~1);
      q.blockDim = dim3(kThreads);
      q.dynamicSmemBytes = Cfg::kSmemBytes;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
      q.attrs = qa;
      q.numAttrs = 1;
      SDPC_CUDA(cudaOccupancyMaxActiveClusters(&max_clusters, kernel, &q));
      if (max_clusters < 1) return set_error(SDPC_ERR_CUDA, "conv_umma: no 2-CTA cluster fits on this device");
    }
    attr_set = true;
  }
  if (!CL) {
    int grid = L.geom.num_tiles < L.num_sms ? L.geom.num_tiles : L.num_sms;
    SDPC_CUDA(launch_k(kernel, dim3(grid), dim3(kThreads), Cfg::kSmemBytes, stream, L.tmap_a, L.tmap_b, L.tmap_a_lo, L.tmap_b_lo, L.tmap_half, L.tmap_half_lo, L.geom, L.epi));
  } else {
    int clusters = L.geom.num_tiles / 2 < max_clusters ? L.geom.num_tiles / 2 : max_clusters;
    if (clusters > L.num_sms / 2) clusters = L.num_sms / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = Cfg::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    SDPC_CUDA(cudaLaunchKernelEx(&cfg, kernel, L.tmap_a, L.tmap_b, L.tmap_a_lo, L.tmap_b_lo, L.tmap_half, L.tmap_half_lo, L.geom, L.epi));
  }
  SDPC_CUDA(cudaGetLastError());
  return SDPC_OK;
}

// swapped layers run as 2-CTA clusters with operand multicast unless SDPC_CLUSTER=0 (A/B switch, read once)
bool conv_umma_cluster() {
  static const bool on = [] { const char* v = getenv("SDPC_CLUSTER"); return !(v && v[0] == '0'); }();
  return on;
}
// the bf16x3 arm (three passes over K with hi / lo operand planes) runs in clusters / CTA pairs like the bf16 arm (measured
// on a B200: outputs bit-identical to single CTAs, 8-view forward 27.4 -> 26.5 ms); SDPC_X3_CLUSTER=0 = single CTAs
bool conv_umma_x3_cluster() {
  static const bool on = [] { const char* v = getenv("SDPC_X3_CLUSTER"); return !(v && v[0] == '0'); }();
  return on;
}
// Cout = 256 layers run swapped as two 128-channel halves unless SDPC_SWAP256=0 (A/B switch, read once)
bool conv_umma_swap256() {
  static const bool on = [] { const char* v = getenv("SDPC_SWAP256"); return !(v && v[0] == '0'); }();
  return on;
}
// Cout = 256 clusters run as CTA pairs on one cta_group::2 MMA unless SDPC_CTA2=0 (A/B switch, read once): measured
// 11.65 -> 11.30 ms per 8-view forward, outputs bit-identical to the multicast clusters
bool conv_umma_cta2() {
  static const bool on = [] { const char* v = getenv("SDPC_CTA2"); return !(v && v[0] == '0'); }();
  return on;
}
// tile_pixels(Cout): pixels per tile the kernel variant for this Cout uses (the host builds the TMA box from it)
int conv_umma_tile_pixels(int Cout) { return (Cout == 128 || conv_umma_swap256()) ? 256 : 128; }
// partial-statistics slots a tile leaves per pixel tile (score_types.cuh, EpiParams::stats)
int conv_umma_stats_parts(int Cout) { return (Cout == 128 || conv_umma_swap256()) ? 2 : 4; }
// rows of the weight TMA box
int conv_umma_weight_rows(int Cout) { return (Cout == 128 || conv_umma_swap256()) ? 128 : Cout; }

int conv_umma_launch(const UmmaConvLaunch& L, cudaStream_t stream) {
  const ConvGeom& g = L.geom;
  const int bk = 128 / L.elem_bytes;
  const bool swapped = g.Cout == 128 || conv_umma_swap256();
  const int tiles = g.N * g.tiles_w * g.tiles_h * (swapped ? g.Cout / 128 : 1);
  if (g.BW * g.BH != conv_umma_tile_pixels(g.Cout) || (1 << g.bw_shift) != g.BW || g.Cin % bk != 0 || g.BW % 32 != 0 ||
      (g.Cout != 128 && g.Cout != 256) || g.num_tiles != tiles)
    return set_error(SDPC_ERR_UNSUPPORTED, "conv_umma: unsupported shape Cin=%d Cout=%d tile=%dx%d tiles=%d", g.Cin, g.Cout,
                     g.BH, g.BW, g.num_tiles);
  const bool cl = L.use_cluster && swapped && (g.passes == 1 || conv_umma_x3_cluster()) && (g.num_tiles % 2) == 0;
  if (L.elem_bytes == 2) {
    if (g.Cout == 128) return cl ? launch_t<__nv_bfloat16, 256, true, 128, true>(L, stream) : launch_t<__nv_bfloat16, 256, true, 128, false>(L, stream);
    if (!swapped) return launch_t<__nv_bfloat16, 256, false, 256, false>(L, stream);
    if (cl && conv_umma_cta2()) return launch_t<__nv_bfloat16, 256, true, 256, true, 2>(L, stream);
    return cl ? launch_t<__nv_bfloat16, 256, true, 256, true>(L, stream) : launch_t<__nv_bfloat16, 256, true, 256, false>(L, stream);
  }
  if (g.Cout == 128) return cl ? launch_t<float, 256, true, 128, true>(L, stream) : launch_t<float, 256, true, 128, false>(L, stream);
  if (!swapped) return launch_t<float, 256, false, 256, false>(L, stream);
  if (cl && conv_umma_cta2()) return launch_t<float, 256, true, 256, true, 2>(L, stream);
  return cl ? launch_t<float, 256, true, 256, true>(L, stream) : launch_t<float, 256, true, 256, false>(L, stream);
}

}  // namespace sdpc
