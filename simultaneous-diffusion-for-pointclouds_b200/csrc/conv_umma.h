// Host-side launch interface of the tcgen05 implicit-GEMM convolution (conv_umma.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "score_types.cuh"

namespace sdpc {

struct UmmaConvLaunch {
  CUtensorMap tmap_a;   // input operand  {C, W+2p, H+2p, N}, box {BK, BW, BH, 1}, 128B swizzle
  CUtensorMap tmap_b;   // weights        {Cin, Cout, taps}, box {BK, conv_umma_weight_rows(Cout), 1}
  CUtensorMap tmap_a_lo, tmap_b_lo;   // residual (lo) planes of the bf16x3 arm (copies of a / b otherwise)
  CUtensorMap tmap_half;   // cluster variant: half-sized box of the operand the two CTAs of a cluster share (see conv_umma.cu)
  CUtensorMap tmap_half_lo;   // the same box on the residual (lo) plane of that operand (bf16x3 arm in clusters; copy of tmap_half otherwise)
  int use_cluster;         // 1: tmap_half is valid and the layer may run as 2-CTA clusters
  ConvGeom geom;
  EpiParams epi;
  int elem_bytes;       // 2 = bf16 (kind::f16), 4 = tf32 (kind::tf32)
  int num_sms;
};

// rank-`rank` tiled tensor map with 128-byte swizzle; dims/box innermost first.
int make_tmap(CUtensorMap* out, void* base, int elem_bytes, int rank, const uint64_t* dims, const uint32_t* box);
int conv_umma_launch(const UmmaConvLaunch& L, cudaStream_t stream);
int conv_umma_tile_pixels(int Cout);   // 256 with swapped operands (Cout = 128, and Cout = 256 as two halves), else 128
int conv_umma_stats_parts(int Cout);   // partial-statistics slots per pixel tile
int conv_umma_weight_rows(int Cout);   // rows of the weight TMA box
bool conv_umma_swap256();
bool conv_umma_cluster();
bool conv_umma_x3_cluster();   // bf16x3 arm in clusters / CTA pairs too (default; SDPC_X3_CLUSTER=0: single CTAs)

}  // namespace sdpc
