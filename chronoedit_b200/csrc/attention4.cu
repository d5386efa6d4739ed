// Self-attention forward, cluster generation: a cluster of TWO CTAs shares every K/V tile through TMA multicast (each CTA
// fetches half of the tile and the copy lands in both CTAs' shared memory: same L2->SM traffic as attention2.cu), but every
// CTA owns ONE 128-query tile and therefore the whole 512 TMEM columns of its SM:
//     S0 [0,128)  S1 [128,256)   two score buffers: S(j+1), S(j+2) are computed while the softmax still works on tile j
//     P0 [256,320) P1 [320,384)  P (packed bf16) has its own columns: nothing aliases, no serial S -> P -> P.V -> S chain
//     O  [384,512)
// The tensor pipe never waits for the softmax of the SAME tile and the softmax never waits for the tensor pipe: the kernel
// is bound by the softmax throughput of one SM (16 ex2/clk -> 1024 clk per 128x128 tile, the same as the MMA time of a tile).
//
// 384 threads: warp 0 TMA producer, warp 1 MMA issuer (one thread), warps 4-11 softmax: TWO threads per query row (warps w
// and w+4 address the same TMEM lanes; each takes 64 of the 128 score columns, 64 of the 128 output columns), partial row
// maxima / sums exchanged through shared memory with a 64-thread named barrier per warp pair.  The softmax loop is software
// pipelined: while tile j is exponentiated (MUFU-bound) the 64 scores of tile j+1 are pulled into a second register array
// and their maximum is folded in between the ex2 instructions.
//
// Replaces F.scaled_dot_product_attention of the self-attention (transformer_chronoedit.py:97-99).
#include <cstdlib>

#include "attention.cuh"

namespace ce {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int NK = 4;  // K ring depth (S is issued up to three tiles ahead of the softmax)
constexpr int NV = 2;  // V ring depth
constexpr int ATTN4_THREADS = 384;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;
constexpr uint32_t HALF_BYTES = TILE_BYTES / 2;    // one 64-wide head-dim half of a tile (128 rows x 128 B)
constexpr uint32_t PART_BYTES = 64 * 64 * 2;       // one TMA box: 64 keys x 64 dims = this CTA's share of a half
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr uint16_t BOTH_CTAS = 0x3;

struct Smem4 {
  static constexpr uint32_t q = 0;
  static constexpr uint32_t k = q + TILE_BYTES;
  static constexpr uint32_t v = k + NK * TILE_BYTES;
  static constexpr uint32_t xchg = v + NV * TILE_BYTES;   // 2 x 2 x 128 floats: partial row maxima (double buffered), reused for the row sums
  static constexpr uint32_t bars = xchg + 2 * 2 * 128 * 4;
  static constexpr uint32_t total = bars + 192;
};
static_assert(Smem4::total <= 227 * 1024, "attention4: shared memory budget");

enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + NK, V_FULL = K_EMPTY + NK, V_EMPTY = V_FULL + NV, S_FULL = V_EMPTY + NV, S_FREE = S_FULL + 2,
       P_FULL = S_FREE + 2, PV_DONE = P_FULL + 2, NUM_BARS4 = PV_DONE + 2 };
static_assert(NUM_BARS4 * 8 + 8 <= 192, "attention4: barrier block");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATTN4_THREADS, 1)
attention4_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem4::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS4);
  float* xchg = reinterpret_cast<float*>(smem + Smem4::xchg);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // == blockIdx.x & 1
  const int q0 = blockIdx.x * BQ;             // may lie beyond Lq for the padding CTA of an odd tile count: it still feeds its peer
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (a.Lk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[chronoedit_b200] attention4: dynamic shared memory not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < NUM_BARS4; ++i) {
      uint32_t count = 1;
      if ((i >= K_EMPTY && i < K_EMPTY + NK) || (i >= V_EMPTY && i < V_EMPTY + NV)) count = 2;   // one commit from each CTA
      if (i >= S_FREE && i < P_FULL + 2) count = 256;                                              // every softmax thread
      mbar_init(&bars[i], count);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before the peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (event-driven)
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[Q_FULL], TILE_BYTES);
        tma_load_3d(smem + Smem4::q, &tma_q, &bars[Q_FULL], h * HD, q0, b);
        tma_load_3d(smem + Smem4::q + HALF_BYTES, &tma_q, &bars[Q_FULL], h * HD + 64, q0, b);
        int k_next = 0, v_next = 0;
        uint64_t t_start = 0;
        uint32_t idle = 0;
        // this CTA fetches keys [crank*64, crank*64+64) of every tile and multicasts them into both CTAs; a ring slot may be
        // refilled once BOTH CTAs' MMAs have consumed it (their commits are multicast to both EMPTY barriers)
        auto load_half_tile = [&](const CUtensorMap* map, uint8_t* slot, uint64_t* full, int tile) {
          mbar_arrive_expect_tx(full, TILE_BYTES);   // own half + the peer's half
          tma_load_3d_mc(slot + crank * PART_BYTES, map, full, h * HD, tile * BKV + (int)crank * 64, b, BOTH_CTAS);
          tma_load_3d_mc(slot + HALF_BYTES + crank * PART_BYTES, map, full, h * HD + 64, tile * BKV + (int)crank * 64, b, BOTH_CTAS);
        };
        while (k_next < n_tiles || v_next < n_tiles) {
          bool progress = false;
          if (k_next < n_tiles) {
            const int st = k_next % NK;
            if (mbar_test_wait(&bars[K_EMPTY + st], ((k_next / NK) & 1) ^ 1)) {
              load_half_tile(&tma_k, smem + Smem4::k + st * TILE_BYTES, &bars[K_FULL + st], k_next);
              ++k_next;
              progress = true;
            }
          }
          if (v_next < n_tiles) {
            const int st = v_next % NV;
            if (mbar_test_wait(&bars[V_EMPTY + st], ((v_next / NV) & 1) ^ 1)) {
              load_half_tile(&tma_v, smem + Smem4::v + st * TILE_BYTES, &bars[V_FULL + st], v_next);
              ++v_next;
              progress = true;
            }
          }
          if (progress) {
            idle = 0;
          } else if ((++idle & 0xFFF) == 0) {
            if (t_start == 0) t_start = global_timer_ns();
            else if (global_timer_ns() - t_start > CE_MBAR_TIMEOUT_NS) {
              printf("[chronoedit_b200] attention4 producer stalled: block=(%d,%d,%d) k=%d v=%d\n", blockIdx.x, blockIdx.y, blockIdx.z, k_next, v_next);
              __trap();
            }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer
      if (lane == 0) {
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, 0);   // Q (K-major, smem) x K^T (K-major, smem)
        constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 1);  // P (TMEM) x V (MN-major, smem)
        const uint32_t q_addr = smem_u32(smem + Smem4::q);
        const uint32_t o_tm = tmem_base + 384;
        auto issue_s = [&](int j) {
          mbar_wait(&bars[K_FULL + j % NK], (j / NK) & 1, 30);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(smem + Smem4::k + (j % NK) * TILE_BYTES);
          const uint32_t d = tmem_base + (j & 1) * 128;
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk) {
            const uint32_t off = (kk >> 2) * HALF_BYTES;
            umma_bf16_ss(d, umma_desc_kmajor_sw128(q_addr + off) + 2 * (kk & 3), umma_desc_kmajor_sw128(k_addr + off) + 2 * (kk & 3), IDESC_S, kk != 0);
          }
          umma_commit(&bars[S_FULL + (j & 1)]);
          umma_commit_mc(&bars[K_EMPTY + j % NK], BOTH_CTAS);
        };
        mbar_wait(&bars[Q_FULL], 0, 1);
        issue_s(0);
        if (n_tiles > 1) issue_s(1);
        if (n_tiles > 2) {
          mbar_wait(&bars[S_FREE + 0], 0, 34);  // S(0) is in the softmax registers
          issue_s(2);
        }
        for (int k = 0; k < n_tiles; ++k) {
          if (k + 3 < n_tiles) {
            // the softmax pulls S(k+1) into registers in the middle of its tile-k step: S(k+3) takes that buffer
            mbar_wait(&bars[S_FREE + ((k + 1) & 1)], ((k + 1) >> 1) & 1, 35);
            issue_s(k + 3);
          }
          mbar_wait(&bars[P_FULL + (k & 1)], (k >> 1) & 1, 40);
          mbar_wait(&bars[V_FULL + k % NV], (k / NV) & 1, 50);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(smem + Smem4::v + (k % NV) * TILE_BYTES);
          const uint32_t p_tm = tmem_base + 256 + (k & 1) * 64;  // packed bf16: 8 columns per K=16 step
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk)
            umma_bf16_ts(o_tm, p_tm + kk * 8, umma_desc_mnmajor_sw128(v_addr + kk * 2048, HALF_BYTES), IDESC_PV, (k | kk) != 0);
          umma_commit(&bars[PV_DONE + (k & 1)]);
          umma_commit_mc(&bars[V_EMPTY + k % NV], BOTH_CTAS);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ---------------------------------------------------------------- softmax: two threads per query row
    const int half = (warp - 4) >> 2;   // which 64 score columns / 64 output columns
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + half * 64;            // + (j&1)*128
    const uint32_t p_tmem = tmem_base + lane_base + 256 + half * 32;      // + (j&1)*64
    const uint32_t o_tmem = tmem_base + lane_base + 384 + half * 64;
    const float sl2 = a.scale * 1.4426950408889634f;
    float m, l = 0.f, alpha = 1.0f;
    bool need = false;
    const bool timed = a.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    long long tc0 = 0;
#define CE_TICK(slot)                      \
  if (timed) {                             \
    const long long _t = clock64();        \
    tacc[slot] += _t - tc0;                \
    tc0 = _t;                              \
  }

    auto mask_tail = [&](uint32_t (&t)[64], int valid) {   // valid = number of in-range keys among this thread's 64
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) t[i] = (i < valid) ? t[i] : 0xff800000u;
      }
    };
    // row maximum over both halves: partial maxima meet in shared memory, one 64-thread barrier per warp pair
    auto row_max = [&](float mine, int buf) {
      xchg[(buf * 2 + half) * 128 + r] = mine;
      named_bar_sync(1 + quad, 64);
      return fmaxf(mine, xchg[(buf * 2 + (half ^ 1)) * 128 + r]);
    };
    auto exp_pair = [&](const uint32_t (&cur)[64], uint32_t (&pk)[32], uint64_t (&sum2)[4], int i, uint64_t sl2_2, uint64_t negm_2) {
      float x0, x1;
      f2_unpack(f2_fma(f2_pack_bits(cur[2 * i], cur[2 * i + 1]), sl2_2, negm_2), x0, x1);
      const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
      sum2[i & 3] = f2_add(sum2[i & 3], f2_pack(p0, p1));
      pk[i] = pack_bf16x2(p0, p1);
    };
    // `cur` holds this thread's 64 scores of tile j; m / alpha / need are already decided for it
    auto step = [&](uint32_t (&cur)[64], uint32_t (&nxt)[64], int j) {
      const bool has_next = j + 1 < n_tiles;
      if (has_next) {
        mbar_wait(&bars[S_FULL + ((j + 1) & 1)], ((j + 1) >> 1) & 1, 60);
        tc_fence_after();
        tmem_ld_32x32(s_tmem + ((j + 1) & 1) * 128, *reinterpret_cast<uint32_t(*)[32]>(&nxt[0]));
        tmem_ld_32x32(s_tmem + ((j + 1) & 1) * 128 + 32, *reinterpret_cast<uint32_t(*)[32]>(&nxt[32]));
      }
      CE_TICK(0)
      const float neg_m = -m;
      const uint64_t sl2_2 = f2_pack(sl2, sl2), negm_2 = f2_pack(neg_m, neg_m);
      uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) exp_pair(cur, pk, sum2, i, sl2_2, negm_2);
      CE_TICK(1)
      float mx8[8];
      if (has_next) {
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(&bars[S_FREE + ((j + 1) & 1)]);  // the buffer may take S(j+3)
        mask_tail(nxt, a.Lk - (j + 1) * BKV - half * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(nxt[i]);
      }
      CE_TICK(2)
#pragma unroll
      for (int i = 16; i < 32; ++i) {
        exp_pair(cur, pk, sum2, i, sl2_2, negm_2);
        if (has_next && i >= 18) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 8 + (i - 18) * 4 + e;
            mx8[c & 7] = fmaxf(mx8[c & 7], __uint_as_float(nxt[c]));
          }
        }
      }
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(sum2[0], sum2[1]), a0, a1);
        f2_unpack(f2_add(sum2[2], sum2[3]), b0, b1);
        l = fmaf(l, alpha, (a0 + a1) + (b0 + b1));  // old sum moves to the new reference (alpha = 1 unless the max jumped)
      }
      CE_TICK(3)
      // P buffer j&1 was last read by P.V(j-2) (one PV_DONE barrier per buffer: a wait never lags its barrier by two phases)
      if (j >= 2) mbar_wait(&bars[PV_DONE + (j & 1)], ((j - 2) >> 1) & 1, 70);
      if (__any_sync(0xffffffffu, need)) {
        mbar_wait(&bars[PV_DONE + ((j - 1) & 1)], ((j - 1) >> 1) & 1, 72);  // O still receives P.V(j-1)
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(o_tmem + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(o_tmem + c * 32, o);
        }
      }
      tc_fence_after();
      tmem_st_32x32(p_tmem + (j & 1) * 64, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
      float mxp = 0.f;
      if (has_next)
        mxp = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[P_FULL + (j & 1)]);
      // reference for tile j+1 (identical in both threads of the row)
      alpha = 1.0f;
      need = false;
      if (has_next) {
        const float mx = sl2 * row_max(mxp, (j + 1) & 1);
        need = mx > m + RESCALE_THRESHOLD;
        if (need) {
          alpha = fast_exp2(m - mx);
          m = mx;
        }
      }
      CE_TICK(4)
    };

    uint32_t sA[64], sB[64];
    mbar_wait(&bars[S_FULL + 0], 0, 60);
    tc_fence_after();
    tmem_ld_32x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&sA[0]));
    tmem_ld_32x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&sA[32]));
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(&bars[S_FREE + 0]);
    mask_tail(sA, a.Lk - half * 64);
    {
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(sA[i]);
#pragma unroll
      for (int i = 8; i < 64; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(sA[i]));
      m = sl2 * row_max(fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7]))), 0);
    }
    if (timed) tc0 = clock64();
    for (int j = 0; j < n_tiles; j += 2) {
      step(sA, sB, j);
      if (j + 1 < n_tiles) step(sB, sA, j + 1);
    }
    if (timed) {
      for (int i = 0; i < 5; ++i) a.timing[i] = tacc[i];
      a.timing[5] = n_tiles;
    }

    // ---- total row sum, normalise and store this thread's 64 output columns
    // (the exchange buffer NOT used by the last row-max exchange: last written two barriers ago)
    const int lb = n_tiles & 1;
    xchg[(lb * 2 + half) * 128 + r] = l;
    named_bar_sync(1 + quad, 64);
    const float inv = 1.0f / (l + xchg[(lb * 2 + (half ^ 1)) * 128 + r]);
    mbar_wait(&bars[PV_DONE + ((n_tiles - 1) & 1)], ((n_tiles - 1) >> 1) & 1, 80);  // commits complete in issue order
    tc_fence_after();
    const int row = q0 + r;
    bf16* orow = a.out + ((size_t)b * a.Lq + row) * a.ldo + h * HD + half * 64;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t o[32];
      tmem_ld_32x32(o_tmem + c * 32, o);
      tmem_ld_wait();
      if (row < a.Lq) {
#pragma unroll
        for (int v4 = 0; v4 < 4; ++v4) {
          float y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) y[i] = __uint_as_float(o[v4 * 8 + i]) * inv;
          *reinterpret_cast<uint4*>(orow + c * 32 + v4 * 8) =
              make_uint4(pack_bf16x2(y[0], y[1]), pack_bf16x2(y[2], y[3]), pack_bf16x2(y[4], y[5]), pack_bf16x2(y[6], y[7]));
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be multicasting into this CTA's shared memory / barriers
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_qkv_tmap4(CUtensorMap* m, const bf16* base, int B, int L, int H, int ld, uint32_t box_rows) {
  uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
  uint32_t box[3] = {64, box_rows, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

}  // namespace

int launch_attention4(const AttnArgs& a, cudaStream_t stream) {
  CE_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0 && a.Lk2 == 0 && a.accumulate == 0, "attention4: single source, no accumulate");
  CE_REQUIRE(a.head_dim == HD, "attention4: only head_dim 128 is built");
  CE_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention4: leading dims % 8");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_tmap4(&tq, a.q, a.B, a.Lq, a.H, a.ldq, 128))) return rc;
  if ((rc = make_qkv_tmap4(&tk, a.k, a.B, a.Lk, a.H, a.ldk, 64))) return rc;
  if ((rc = make_qkv_tmap4(&tv, a.v, a.B, a.Lk, a.H, a.ldv, 64))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    CE_CHECK_CUDA(cudaFuncSetAttribute(attention4_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem4::total));
    attr_set = true;
  }
  const int q_tiles = (a.Lq + BQ - 1) / BQ;
  dim3 grid(2 * ((q_tiles + 1) / 2), a.H, a.B);   // whole clusters: an odd tile count gets one padding CTA
  attention4_fwd_kernel<<<grid, ATTN4_THREADS, Smem4::total, stream>>>(tq, tk, tv, a);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
