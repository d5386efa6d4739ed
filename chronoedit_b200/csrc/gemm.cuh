// Interface of the tcgen05 GEMM used by the DiT path (gemm.cu).
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace ce {

// Fused epilogues.  "y" is the Linear output rounded to bf16 exactly as nn.Linear returns it in the
// reference's bf16 path; every variant keeps the reference's rounding points.
enum GemmEpilogue : int {
  EPI_BIAS = 0,             // out = bf16(acc + bias)
  EPI_BIAS_GELU_TANH = 1,   // out = bf16(gelu_tanh(float(bf16(acc + bias))))          ffn.net.0  (transformer_chronoedit.py:292)
  EPI_BIAS_GELU_ERF = 2,    // out = bf16(gelu_erf (float(bf16(acc + bias))))          image_embedder.ff.net.0 (:111-123)
  EPI_BIAS_GATE_RESID = 3,  // out = bf16(float(resid) + float(bf16(acc + bias)) * gate[batch, col])   (:281, :293)
  EPI_BIAS_RESID = 4,       // out = bf16(float(resid) + float(bf16(acc + bias)))                      (:286)
};

struct GemmArgs {
  int M = 0, N = 0, K = 0;
  bf16* out = nullptr;        // [M, ldo] bf16 (may be null when out_f32 is given)
  int ldo = 0;
  float* out_f32 = nullptr;   // optional [M, ld_f32] fp32 copy of (acc + bias) BEFORE any rounding (parity tests, fp32 validation mode)
  int ld_f32 = 0;             // leading dimension of out_f32 in elements (0 = N)
  const bf16* bias = nullptr; // [N] or null
  const bf16* bias_row = nullptr;  // [M] or null: added per output ROW (used for V^T = W_v x^T + b_v in the VAE attention)
  int epi = EPI_BIAS;
  const bf16* resid = nullptr;  // [M, ldr]
  int ldr = 0;
  const float* gate = nullptr;  // [batches, gate_stride] fp32, row -> batch = row / rows_per_batch
  int gate_stride = 0;
  int rows_per_batch = 1;
  // Row statistics of the STORED bf16 output, one float2 per (row, N tile) at stats_out[row * stats_ld + n_tile]:
  //   stats_mode 1: (mean, M2 = sum (y - mean)^2) of the tile's columns      -> FP32LayerNorm of the next operator
  //   stats_mode 2: (sum y^2, 0)                                             -> RMSNorm across heads
  // so that the row kernel that follows is a pure streaming pass (no reduction, no barrier); tiles are merged there in a fixed
  // order (deterministic).  N tile width = gemm_tile_n(N).
  float2* stats_out = nullptr;
  int stats_ld = 0;
  int stats_mode = 0;
  int group_m = 0;              // tile rasterisation: M-tiles per group (L2 reuse); 0 = chosen by launch_gemm_bf16
};

// N tile width launch_gemm_bf16 uses for a given N (the granularity of GemmArgs::stats_out)
inline int gemm_tile_n(int N) { return N >= 256 ? 256 : (N >= 128 ? 128 : 64); }

// out[M,N] = epilogue(A[M,K] (row-major, lda) x W[N,K]^T (row-major = nn.Linear weight, ldw)).
int launch_gemm_bf16(const bf16* A, int lda, const bf16* W, int ldw, const GemmArgs& g, cudaStream_t stream);

}  // namespace ce
