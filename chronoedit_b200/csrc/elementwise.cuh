// Interface of the HBM-bound row kernels of the DiT path (elementwise.cu).
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace ce {

// y[row, :] = bf16( LN_fp32(x[row, :]) * (1 + scale[b, :]) + shift[b, :] )     (b = row / rows_per_batch)
//   scale/shift: fp32 [batches, mod_stride] or null (then the optional affine weight/bias, fp32 [D], is applied)
// FP32LayerNorm + adaLN modulate of transformer_chronoedit.py:279, 289, 460 and the affine norm2 of :284.
//   scale_is_1p: the `scale` table already holds (1 + scale) (ce_dit stores its modulation tables that way)
int launch_layernorm(const bf16* x, int ldx, bf16* y, int ldy, int rows, int D, float eps, const float* scale,
                     const float* shift, int mod_stride, int rows_per_batch, const float* weight, const float* bias,
                     cudaStream_t stream, int scale_is_1p = 0);

// In place on x[rows, D] (leading dim ldx): diffusers RMSNorm across all heads (fp32 variance over D,
// y = bf16(bf16(x * rstd) * w)), then optionally interleaved-pair RoPE with cos/sin tables [L, hd/2] fp32
// (token = row % L).  transformer_chronoedit.py:62-65 and :71-79.
int launch_rmsnorm_rope(bf16* x, int ldx, int rows, int D, float eps, const bf16* weight, const float* rope_cos,
                        const float* rope_sin, int L, int head_dim, cudaStream_t stream);

// The same two operators when the row statistics come from the epilogue of the GEMM that produced x (GemmArgs::stats_out,
// per-(row, N tile) partials): streaming kernels without a reduction or a barrier.  `tiles` partials of width tile_n cover [0, D).
// The RMSNorm variant handles `nmat` (1 or 2) matrices lying side by side in a row (q | k of the fused QKV output): matrix m starts
// at column m * D, uses weight m and partials [m * tiles, (m + 1) * tiles).
int launch_layernorm_stats(const bf16* x, int ldx, bf16* y, int ldy, int rows, int D, float eps, const float* scale, const float* shift,
                           int mod_stride, int rows_per_batch, const float* weight, const float* bias, int scale_is_1p, const float2* stats,
                           int stats_ld, int tiles, int tile_n, cudaStream_t stream);
// Sequence-parallel scatter of the normalised q | k (csrc/seqpar.cu): rows are the LOCAL tokens of this rank (rows_per_batch each,
// global token = tok0 + local index), RoPE uses the global token, and instead of being written back in place the vector of head h
// goes to the rank that owns head h:  dst[mat][h / heads_per_rank] + ((b * L_total + token) * (heads_per_rank * head_dim)
// + (h % heads_per_rank) * head_dim + offset inside the head).
struct SpScatter {
  bf16* dst[2][8];
  int world = 0;            // 0 = no scatter (in place)
  int heads_per_rank = 0;
  int L_total = 0;          // tokens per sample over all ranks
  int rows_per_batch = 0;   // local tokens per sample
  int tok0 = 0;             // first global token of this rank
};
int launch_rmsnorm_rope_stats(bf16* x, int ldx, int rows, int D, float eps, const bf16* weight0, const bf16* weight1, int nmat,
                              const float* rope_cos, const float* rope_sin, int L, int head_dim, const float2* stats, int stats_ld, int tiles,
                              cudaStream_t stream, const SpScatter* sp = nullptr);

// patches[(b,f,i,j), c*4 + dh*2 + dw] = x[b, c, f, 2i+dh, 2j+dw]   (im2row for the k=s=(1,2,2) patch-embedding conv, :368,429-430)
int launch_patchify(const bf16* x, bf16* patches, int B, int C, int T, int H, int W, cudaStream_t stream);
// out[b, c, f, 2i+dh, 2j+dw] = y[(b,f,i,j), (dh*2+dw)*C + c]       (unpatchify permute of :463-467)
int launch_unpatchify(const bf16* y, int ldy, bf16* out, int B, int C, int T, int H, int W, cudaStream_t stream);

// out[b, n] = act( sum_k x[b,k] * W[n,k] + bias[n] );  tiny-M Linear (time embedder / time_proj, :155-159).
//   w_is_bf16: weights/bias bf16 (else fp32); x is fp32; out fp32 (out_f32) and/or bf16 (out_bf16).
//   in_silu_bf16: the input is first mapped through bf16(SiLU(x)) (act_fn applied to the bf16 temb tensor).
//   act: 0 none, 1 SiLU on the output.  When out_bf16 is given the Linear result is rounded to bf16 first (and
//   out_f32, if also given, receives that rounded value).
int launch_small_linear(const float* x, int K, const void* W, const void* bias, int w_is_bf16, int N, int B,
                        int act, int in_silu_bf16, float* out_f32, bf16* out_bf16, cudaStream_t stream);

// emb[b, :] = [cos(t_b * f_i) | sin(t_b * f_i)], f_i = exp(-ln(10000) * i / half)   (diffusers Timesteps, flip_sin_to_cos)
int launch_timestep_sinusoid(const float* t, float* emb, int B, int dim, cudaStream_t stream);

// dst[b, :] = table[b % table_rows, :] + src[b, :] (fp32): scale_shift_table + temb (:274-276, :451)
//   plus_one_mask: bit c set -> chunk c is stored as 1 + value (the adaLN "scale" chunks)
int launch_add_table(const float* table, int table_rows, const bf16* src, int src_ld, float* dst, int B, int n,
                     int chunks, cudaStream_t stream, unsigned plus_one_mask = 0);

}  // namespace ce
