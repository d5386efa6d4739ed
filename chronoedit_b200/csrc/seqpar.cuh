// Sequence parallelism for single-edit latency: the tokens of ONE edit are split over the ranks of
// a node, every rank holds all weights, and the only exchange is around the self-attention (heads <-> tokens, "Ulysses" layout,
// reference: chronoedit_diffsynth/wan_video_new_chronoedit.py:330-355, 1448-1498).  Here the exchange is not a collective call:
// the kernels that PRODUCE the data store it straight into the consuming rank's memory over NVLink (peer pointers obtained with
// CUDA IPC) -- the RMSNorm+RoPE kernel scatters q | k by head, a copy kernel scatters v, the attention epilogue scatters the
// output rows back to the ranks that own the tokens -- and a flag barrier in peer memory orders the phases.
#pragma once
#include "host_common.h"
#include "ptx.cuh"

namespace ce {

struct SeqPar {
  int rank = 0, world = 1;
  uint8_t* region[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // same layout on every rank
  int64_t region_bytes = 0;
  uint32_t epoch = 0;
};

struct SpLayout {   // byte offsets inside a rank's region
  int64_t q, k, v;      // gathered q / k / v of the heads this rank owns: [B, L, D / world] bf16 each
  int64_t attn;         // attention output of the tokens this rank owns: [B * L / world, D] bf16
  int64_t yout;         // full head output [B * L, out_channels * 4] bf16 (every rank gets all of it)
  int64_t flags;        // uint32 [8 * 32]
  int64_t bytes;
};
SpLayout sp_layout(int64_t B, int64_t L, int64_t D, int64_t No, int world);

// all ranks: everything this rank issued so far on `stream` (incl. its peer stores) is complete and visible before any rank continues
int launch_sp_barrier(SeqPar& sp, const SpLayout& lay, cudaStream_t stream);
// dst[r] + ((b * dst_rows_per_batch + dst_row0 + i) * cols_per_rank + c % cols_per_rank) = src[(b * rows_per_batch + i) * ld_src + c],
// r = c / cols_per_rank   (v scatter: columns to the rank that owns the head)
int launch_sp_scatter_cols(const bf16* src, int ld_src, int B, int rows_per_batch, int ncols, bf16* const* dst, int world, int cols_per_rank,
                           int dst_rows_per_batch, int dst_row0, cudaStream_t stream);
// every rank r: dst[r] + ((b * dst_rows_per_batch + dst_row0 + i) * ncols + c) = src[(b * rows_per_batch + i) * ncols + c]   (all-gather of rows)
int launch_sp_broadcast_rows(const bf16* src, int B, int rows_per_batch, int ncols, bf16* const* dst, int world, int dst_rows_per_batch,
                             int dst_row0, cudaStream_t stream);
// patchify of a token range: patches[(b, i), k] for global tokens tok0 .. tok0 + ntok - 1 of every sample
int launch_patchify_range(const bf16* x, bf16* patches, int B, int C, int T, int H, int W, int tok0, int ntok, cudaStream_t stream);

}  // namespace ce
