// 2-CTA (cta_group::2) variant of the bf16 GEMM: a cluster of two CTAs on one TPC computes a 256 x 256 output tile.
//
// Each CTA TMA-loads its own 128 rows of A and HALF of the W tile (128 of the 256 N rows) per k-block — 32 KB per stage
// per CTA instead of 48 KB, so six stages fit and L2->SM and shared-memory operand traffic drop by a third — and ONE
// thread of the leader CTA issues tcgen05.mma.cta_group::2 (M256 x N256 x K16), which reads both CTAs' shared memory and
// writes each CTA's 128 accumulator rows into that CTA's own TMEM.  Completion of a stage is tracked on the LEADER's
// full barrier (both CTAs' TMA loads credit it); tcgen05.commit multicasts the "slot free" / "accumulator ready" arrivals
// to both CTAs; the epilogue warps of both CTAs release the accumulator stage on the leader's barrier.
// Same epilogues, tile rasterisation and launch interface as gemm.cu (which remains the path for small problems).
#include <cstdlib>

#include "gemm.cuh"

namespace ce {

namespace {

constexpr int BM = 128;       // rows per CTA
constexpr int BN = 256;       // N tile of the pair (each CTA stages BN/2 rows of W)
constexpr int BK = 64;
constexpr int STAGES = 6;
constexpr int THREADS = 192;
constexpr uint32_t A_BYTES = BM * BK * 2;
constexpr uint32_t B_BYTES = (BN / 2) * BK * 2;
constexpr uint32_t TMEM_COLS = 2 * BN;

__device__ __forceinline__ void tile_coords2(int t, int tiles_m, int tiles_n, int group_m, int& mb, int& nb) {
  const int per_group = group_m * tiles_n;
  const int g = t / per_group;
  const int first_m = g * group_m;
  const int gsize = min(tiles_m - first_m, group_m);
  const int r = t - g * per_group;
  mb = first_m + r % gsize;
  nb = r / gsize;
}

__device__ __forceinline__ float gelu_tanh_f2(float x) {
  const float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(inner));
}
__device__ __forceinline__ float gelu_erf_f2(float x) { return 0.5f * x * (1.0f + erff(x * 0.7071067811865476f)); }

// ELECT: the single producer / issuer lane is chosen with elect.sync instead of `lane == 0` -- under elect.sync the compiler knows
// the region runs in one thread and feeds the uniform-register operands of UTMALDG / UTCHMMA directly; under `lane == 0` it wraps
// every such instruction in a serialisation loop (ELECT / BRA.U.ANY), ~10 extra instructions per MMA (A/B: CE_GEMM_ELECT=0).
template <bool ELECT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(THREADS, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, GemmArgs g) {
  constexpr uint32_t IDESC = umma_idesc_bf16(2 * BM, BN, 0);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * (A_BYTES + B_BYTES));
  uint64_t* full = bars;                    // used on the leader only (count 2: leader expect_tx + peer arrive)
  uint64_t* empty = bars + STAGES;          // per CTA, signalled by the leader's multicast commit
  uint64_t* tfull = bars + 2 * STAGES;      // per CTA, multicast commit
  uint64_t* tempty = bars + 2 * STAGES + 2; // leader only: 8 arrivals (4 epilogue warps x 2 CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int tiles_m = (g.M + 2 * BM - 1) / (2 * BM);
  const int tiles_n = (g.N + BN - 1) / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int kblocks = (g.K + BK - 1) / BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 2);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
  }
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA credit
  if (warp == 1) tmem_alloc_2sm(tmem_slot, TMEM_COLS);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (one lane in EACH CTA)
    if (ELECT ? elect_one_sync() : lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mb, nb;
        tile_coords2(t, tiles_m, tiles_n, g.group_m, mb, nb);
        const int m0 = mb * 2 * BM + (int)rank * BM;
        const int n0 = nb * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1, 100 + stage);
          if (leader) mbar_arrive_expect_tx(&full[stage], 2 * (A_BYTES + B_BYTES));
          tma_load_2d_2sm(sA + stage * A_BYTES, &tma_a, &full[stage], kb * BK, m0);
          tma_load_2d_2sm(sB + stage * B_BYTES, &tma_b, &full[stage], kb * BK, n0);
          if (!leader) mbar_arrive_remote(&full[stage], 0);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (leader && (ELECT ? elect_one_sync() : lane == 0)) {
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1, 200 + acc);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(&full[stage], phase, 300 + stage);
          tc_fence_after();
          const uint64_t da = umma_desc_kmajor_sw128(smem_u32(sA + stage * A_BYTES));
          const uint64_t db = umma_desc_kmajor_sw128(smem_u32(sB + stage * B_BYTES));
          if (ELECT) {
            umma_bf16_ss_2sm_x4(d_tmem, da, db, IDESC, kb != 0);   // the four K16 steps of the k-block, one asm statement
          } else {
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) umma_bf16_ss_2sm(d_tmem, da + 2 * k, db + 2 * k, IDESC, (kb | k) != 0);
          }
          umma_commit_2sm(&empty[stage], 3);  // both CTAs' producers
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
        umma_commit_2sm(&tfull[acc], 3);  // both CTAs' epilogues
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5 of each CTA: its own 128 rows)
    const int q = warp & 3;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters) {
      int mb, nb;
      tile_coords2(t, tiles_m, tiles_n, g.group_m, mb, nb);
      mbar_wait(&tfull[acc], acc_phase, 400 + acc);
      tc_fence_after();
      const int row = mb * 2 * BM + (int)rank * BM + q * 32 + lane;
      const bool row_ok = row < g.M;
      const int batch = row_ok ? row / g.rows_per_batch : 0;
      const uint32_t t_row = tmem_base + (uint32_t(q * 32) << 16) + acc * BN;
      float st_pivot = 0.f, st_s = 0.f, st_ss = 0.f;   // row statistics of this tile (GemmArgs::stats_out)
      int st_n = 0;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int n0 = nb * BN + c * 32;
        if (n0 >= g.N) break;
        uint32_t r[32];
        tmem_ld_32x32(t_row + c * 32, r);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int v = 0; v < 4; ++v) {
            const int n = n0 + v * 8;
            if (n >= g.N) break;
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = __uint_as_float(r[v * 8 + j]);
            if (g.bias) {
              const uint4 bv = *reinterpret_cast<const uint4*>(g.bias + n);
              const float2 b0 = unpack_bf16x2(bv.x), b1 = unpack_bf16x2(bv.y), b2 = unpack_bf16x2(bv.z), b3 = unpack_bf16x2(bv.w);
              y[0] += b0.x; y[1] += b0.y; y[2] += b1.x; y[3] += b1.y;
              y[4] += b2.x; y[5] += b2.y; y[6] += b3.x; y[7] += b3.y;
            }
            if (g.bias_row) {
              const float br = __bfloat162float(g.bias_row[row]);
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] += br;
            }
            if (g.out_f32) {
              float4* o = reinterpret_cast<float4*>(g.out_f32 + (size_t)row * (g.ld_f32 ? g.ld_f32 : g.N) + n);
              o[0] = make_float4(y[0], y[1], y[2], y[3]);
              o[1] = make_float4(y[4], y[5], y[6], y[7]);
              if (!g.out) continue;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = bf16_round(y[j]);
            if (g.epi == EPI_BIAS_GELU_TANH) {
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = gelu_tanh_f2(y[j]);
            } else if (g.epi == EPI_BIAS_GELU_ERF) {
#pragma unroll
              for (int j = 0; j < 8; ++j) y[j] = gelu_erf_f2(y[j]);
            } else if (g.epi == EPI_BIAS_GATE_RESID || g.epi == EPI_BIAS_RESID) {
              const uint4 xv = *reinterpret_cast<const uint4*>(g.resid + (size_t)row * g.ldr + n);
              const float2 x0 = unpack_bf16x2(xv.x), x1 = unpack_bf16x2(xv.y), x2 = unpack_bf16x2(xv.z), x3 = unpack_bf16x2(xv.w);
              const float x[8] = {x0.x, x0.y, x1.x, x1.y, x2.x, x2.y, x3.x, x3.y};
              if (g.epi == EPI_BIAS_GATE_RESID) {
                const float4* gp = reinterpret_cast<const float4*>(g.gate + (size_t)batch * g.gate_stride + n);
                const float4 g0 = gp[0], g1 = gp[1];
                const float gt[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = x[j] + y[j] * gt[j];
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) y[j] = x[j] + y[j];
              }
            }
            if (g.stats_mode) {   // statistics of the values as stored (bf16), pivot-shifted inside the tile
              float z[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) z[j] = bf16_round(y[j]);
              if (st_n == 0) st_pivot = g.stats_mode == 1 ? z[0] : 0.f;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float d = z[j] - st_pivot;
                st_s += d;
                st_ss = fmaf(d, d, st_ss);
              }
              st_n += 8;
            }
            *reinterpret_cast<uint4*>(g.out + (size_t)row * g.ldo + n) =
                make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
          }
        }
      }
      if (g.stats_mode && row_ok && st_n > 0) {
        float2 pr;
        if (g.stats_mode == 1) {   // (mean, M2) of this tile's st_n columns
          const float md = st_s / (float)st_n;
          pr = make_float2(st_pivot + md, fmaxf(st_ss - st_s * md, 0.f));
        } else {
          pr = make_float2(st_ss, 0.f);
        }
        g.stats_out[(size_t)row * g.stats_ld + nb] = pr;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tempty[acc]);
        else mbar_arrive_remote(&tempty[acc], 0);
      }
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair still touches its memory / barriers
  if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
}

}  // namespace

int launch_gemm_bf16_2cta(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, cudaStream_t stream) {
  constexpr size_t smem = (size_t)STAGES * (A_BYTES + B_BYTES) + 1024 + 256;
  static const bool elect = [] {
    const char* e = getenv("CE_GEMM_ELECT");
    return !(e && e[0] == '0');
  }();
  const int tiles = ((g.M + 2 * BM - 1) / (2 * BM)) * ((g.N + BN - 1) / BN);
  int clusters = device_sm_count() / 2;
  if (tiles < clusters) clusters = tiles;
  if (elect) {
    CE_ENSURE_SMEM(gemm_bf16_2cta_kernel<true>, smem);
    gemm_bf16_2cta_kernel<true><<<2 * clusters, THREADS, smem, stream>>>(ta, tb, g);
  } else {
    CE_ENSURE_SMEM(gemm_bf16_2cta_kernel<false>, smem);
    gemm_bf16_2cta_kernel<false><<<2 * clusters, THREADS, smem, stream>>>(ta, tb, g);
  }
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
