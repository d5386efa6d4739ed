// Self-attention forward, third generation.  Two 128-query tiles per CTA share every K/V tile (as attention2.cu), but
//   * key tiles are 64 wide, so one query tile needs only 64 TMEM columns for S and P gets its OWN columns
//     (two 32-column buffers): S(j+1) no longer has to wait for P.V(j) -- it is issued the moment the softmax group has
//     pulled S(j) into registers, and is ready long before the group finishes exp / P(j);
//   * each query tile has its own MMA issuer thread (warp 1 / warp 2): the events of one chain strictly alternate
//     (S drained -> issue next S;  P written -> issue P.V), so plain blocking waits in program order suffice.
// The softmax groups therefore run back to back (MUFU-bound: 2 x 64 x 128 exp2 per 64-key step = the MMA time of that step).
//   TMEM:  q0: S [0,64) P [64,96) [96,128)   q1: S [128,192) P [192,224) [224,256)   O0 [256,384)   O1 [384,512)
//   smem:  Q 2 x 32 KB | K ring 6 x 16 KB | V ring 4 x 16 KB
#include "attention.cuh"

namespace ce {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 64;
constexpr int NK = 6;  // K ring depth
constexpr int NV = 4;  // V ring depth
constexpr int ATTN3_THREADS = 384;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;   // Q tile
constexpr uint32_t HALF_BYTES = TILE_BYTES / 2;
constexpr uint32_t KV_BYTES = BKV * 128 * 2;       // K / V tile (64 keys)
constexpr uint32_t KV_HALF = KV_BYTES / 2;
constexpr float RESCALE_THRESHOLD = 8.0f;

struct Smem3 {
  static constexpr uint32_t q = 0;                         // 2 tiles
  static constexpr uint32_t k = q + 2 * TILE_BYTES;        // NK tiles
  static constexpr uint32_t v = k + NK * KV_BYTES;         // NV tiles
  static constexpr uint32_t bars = v + NV * KV_BYTES;
  static constexpr uint32_t total = bars + 512;
};

enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + NK, V_FULL = K_EMPTY + NK, V_EMPTY = V_FULL + NV, S_FULL = V_EMPTY + NV, S_FREE = S_FULL + 2,
       P_FULL = S_FREE + 2 /* [q*2 + buf] */, PV_DONE = P_FULL + 4 /* [q*2 + buf] */, STAGGER = PV_DONE + 4, NUM_BARS3 = STAGGER + 1 };

__global__ void __launch_bounds__(ATTN3_THREADS, 1)
attention3_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem3::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS3);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (a.Lk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[chronoedit_b200] attention3: dynamic shared memory not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < NUM_BARS3; ++i) mbar_init(&bars[i], ((i >= S_FREE && i < P_FULL + 4) || i == STAGGER) ? 128 : ((i >= K_EMPTY && i < K_EMPTY + NK) || (i >= V_EMPTY && i < V_EMPTY + NV) ? 2 : 1));
    fence_mbar_init();
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (event-driven)
      if (lane == 0) {
        mbar_arrive_expect_tx(&bars[Q_FULL], 2 * TILE_BYTES);
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
          tma_load_3d(smem + Smem3::q + qt * TILE_BYTES, &tma_q, &bars[Q_FULL], h * HD, q0 + qt * BQ, b);
          tma_load_3d(smem + Smem3::q + qt * TILE_BYTES + HALF_BYTES, &tma_q, &bars[Q_FULL], h * HD + 64, q0 + qt * BQ, b);
        }
        int k_next = 0, v_next = 0;
        uint64_t t_start = 0;
        uint32_t idle = 0;
        while (k_next < n_tiles || v_next < n_tiles) {
          bool progress = false;
          if (k_next < n_tiles) {
            const int st = k_next % NK;
            if (mbar_test_wait(&bars[K_EMPTY + st], ((k_next / NK) & 1) ^ 1)) {
              uint8_t* ks = smem + Smem3::k + st * KV_BYTES;
              mbar_arrive_expect_tx(&bars[K_FULL + st], KV_BYTES);
              tma_load_3d(ks, &tma_k, &bars[K_FULL + st], h * HD, k_next * BKV, b);
              tma_load_3d(ks + KV_HALF, &tma_k, &bars[K_FULL + st], h * HD + 64, k_next * BKV, b);
              ++k_next;
              progress = true;
            }
          }
          if (v_next < n_tiles) {
            const int st = v_next % NV;
            if (mbar_test_wait(&bars[V_EMPTY + st], ((v_next / NV) & 1) ^ 1)) {
              uint8_t* vs = smem + Smem3::v + st * KV_BYTES;
              mbar_arrive_expect_tx(&bars[V_FULL + st], KV_BYTES);
              tma_load_3d(vs, &tma_v, &bars[V_FULL + st], h * HD, v_next * BKV, b);
              tma_load_3d(vs + KV_HALF, &tma_v, &bars[V_FULL + st], h * HD + 64, v_next * BKV, b);
              ++v_next;
              progress = true;
            }
          }
          if (progress) {
            idle = 0;
          } else if ((++idle & 0xFFF) == 0) {
            if (t_start == 0) t_start = global_timer_ns();
            else if (global_timer_ns() - t_start > CE_MBAR_TIMEOUT_NS) {
              printf("[chronoedit_b200] attention3 producer stalled: block=(%d,%d,%d) k=%d v=%d\n", blockIdx.x, blockIdx.y, blockIdx.z, k_next, v_next);
              __trap();
            }
          }
        }
      }
    } else if (warp == 1 || warp == 2) {
      // ---------------------------------------------------------------- MMA issuers: warp 1 -> query tile 0, warp 2 -> query tile 1
      if (lane == 0) {
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, BKV, 0);   // Q (K-major, smem) x K^T (K-major, smem): N = 64 keys
        constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 1);  // P (TMEM) x V (MN-major, smem): K = 64 keys
        const int qt = warp - 1;
        const uint32_t q_addr = smem_u32(smem + Smem3::q + qt * TILE_BYTES);
        const uint32_t s_tm = tmem_base + qt * 128;
        const uint32_t o_tm = tmem_base + 256 + qt * 128;
        auto issue_s = [&](int j) {
          mbar_wait(&bars[K_FULL + j % NK], (j / NK) & 1, 30 + qt);
          tc_fence_after();
          const uint32_t k_addr = smem_u32(smem + Smem3::k + (j % NK) * KV_BYTES);
#pragma unroll
          for (int kk = 0; kk < HD / 16; ++kk)
            umma_bf16_ss(s_tm, umma_desc_kmajor_sw128(q_addr + (kk >> 2) * HALF_BYTES) + 2 * (kk & 3),
                         umma_desc_kmajor_sw128(k_addr + (kk >> 2) * KV_HALF) + 2 * (kk & 3), IDESC_S, kk != 0);
          umma_commit(&bars[S_FULL + qt]);
          umma_commit(&bars[K_EMPTY + j % NK]);  // one of the two arrivals (both query tiles consume K_j)
        };
        mbar_wait(&bars[Q_FULL], 0, 1);
        if (qt == 1) mbar_wait(&bars[STAGGER], 0, 2);  // start behind query tile 0 so the two exp phases interleave
        issue_s(0);
        if (n_tiles > 1) {
          mbar_wait(&bars[S_FREE + qt], 0, 35 + qt);  // S(0) is in the group's registers
          issue_s(1);
        }
        for (int j = 0; j < n_tiles; ++j) {
          if (j + 2 < n_tiles) {
            // the group pulls S(j+1) into registers in the middle of its tile-j step: S(j+2) is then ready when that step ends
            mbar_wait(&bars[S_FREE + qt], (j + 1) & 1, 35 + qt);
            issue_s(j + 2);
          }
          mbar_wait(&bars[P_FULL + qt * 2 + (j & 1)], (j >> 1) & 1, 40 + qt);
          mbar_wait(&bars[V_FULL + j % NV], (j / NV) & 1, 50 + qt);
          tc_fence_after();
          const uint32_t v_addr = smem_u32(smem + Smem3::v + (j % NV) * KV_BYTES);
          const uint32_t p_tm = s_tm + 64 + (j & 1) * 32;  // packed bf16: 8 columns per K=16 step
#pragma unroll
          for (int kk = 0; kk < BKV / 16; ++kk)
            umma_bf16_ts(o_tm, p_tm + kk * 8, umma_desc_mnmajor_sw128(v_addr + kk * 2048, KV_HALF), IDESC_PV, (j | kk) != 0);
          umma_commit(&bars[PV_DONE + qt * 2 + (j & 1)]);
          umma_commit(&bars[V_EMPTY + j % NV]);
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
    // ---------------------------------------------------------------- softmax groups (one per query tile)
    const int qt = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + qt * 128;
    const uint32_t o_tmem = tmem_base + lane_base + 256 + qt * 128;
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

    // Software pipeline: while tile j is exponentiated (MUFU-bound), S(j+1) is pulled into the other register array and
    // its row maximum is folded in between the exp2 instructions.  `cur` holds S(j) with m / alpha / need already decided.
    auto mask_tail = [&](uint32_t (&t)[64], int valid) {
      if (valid < BKV) {
#pragma unroll
        for (int i = 0; i < 64; ++i) t[i] = (i < valid) ? t[i] : 0xff800000u;
      }
    };
    auto exp_pair = [&](const uint32_t (&cur)[64], uint32_t (&pk)[32], uint64_t (&sum2)[4], int i, uint64_t sl2_2, uint64_t negm_2) {
      float x0, x1;
      f2_unpack(f2_fma(f2_pack_bits(cur[2 * i], cur[2 * i + 1]), sl2_2, negm_2), x0, x1);
      const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
      sum2[i & 3] = f2_add(sum2[i & 3], f2_pack(p0, p1));
      pk[i] = pack_bf16x2(p0, p1);
    };
    auto step = [&](uint32_t (&cur)[64], uint32_t (&nxt)[64], int j) {
      const bool has_next = j + 1 < n_tiles;
      if (has_next) {
        mbar_wait(&bars[S_FULL + qt], (j + 1) & 1, 60 + qt);
        tc_fence_after();
        tmem_ld_32x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&nxt[0]));
        tmem_ld_32x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&nxt[32]));
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
        mbar_arrive(&bars[S_FREE + qt]);  // S columns may be overwritten by S(j+2)
        mask_tail(nxt, a.Lk - (j + 1) * BKV);
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
      // P buffer j&1 was last read by P.V(j-2).  One PV_DONE barrier per buffer, so a wait never lags its barrier by
      // more than one phase (P.V(j) cannot be issued before this thread publishes P(j)).
      if (j >= 2) mbar_wait(&bars[PV_DONE + qt * 2 + (j & 1)], ((j - 2) >> 1) & 1, 70 + qt);
      if (__any_sync(0xffffffffu, need)) {
        mbar_wait(&bars[PV_DONE + qt * 2 + ((j - 1) & 1)], ((j - 1) >> 1) & 1, 72 + qt);  // O still receives P.V(j-1)
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t o[32];
          tmem_ld_32x32(o_tmem + c * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st_32x32(o_tmem + c * 32, o);
        }
      }
      tc_fence_after();
      tmem_st_32x32(s_tmem + 64 + (j & 1) * 32, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
      // decide the reference for tile j+1 while the store is in flight
      alpha = 1.0f;
      need = false;
      if (has_next) {
        const float mx = sl2 * fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        need = mx > m + RESCALE_THRESHOLD;
        if (need) {
          alpha = fast_exp2(m - mx);
          m = mx;
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[P_FULL + qt * 2 + (j & 1)]);
      CE_TICK(4)
    };

    uint32_t sA[64], sB[64];
    mbar_wait(&bars[S_FULL + qt], 0, 60 + qt);
    tc_fence_after();
    tmem_ld_32x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&sA[0]));
    tmem_ld_32x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&sA[32]));
    tmem_ld_wait();
    tc_fence_before();
    mbar_arrive(&bars[S_FREE + qt]);
    if (qt == 0) mbar_arrive(&bars[STAGGER]);  // query tile 1 starts half a period behind (the two groups share the MUFU)
    mask_tail(sA, a.Lk);
    {
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(sA[i]);
#pragma unroll
      for (int i = 8; i < 64; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(sA[i]));
      m = sl2 * fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
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

    // ---- normalise and store this query tile
    // the last P.V (commits complete in issue order, so every earlier one is done too)
    mbar_wait(&bars[PV_DONE + qt * 2 + ((n_tiles - 1) & 1)], ((n_tiles - 1) >> 1) & 1, 80 + qt);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int row = q0 + qt * BQ + r;
    bf16* orow = a.out + ((size_t)b * a.Lq + row) * a.ldo + h * HD;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
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
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_qkv_tmap3(CUtensorMap* m, const bf16* base, int B, int L, int H, int ld, uint32_t box_rows) {
  uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
  uint32_t box[3] = {64, box_rows, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

}  // namespace

int launch_attention3(const AttnArgs& a, cudaStream_t stream) {
  CE_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0 && a.Lk2 == 0 && a.accumulate == 0, "attention3: single source, no accumulate");
  CE_REQUIRE(a.head_dim == HD, "attention3: only head_dim 128 is built");
  CE_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention3: leading dims % 8");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_tmap3(&tq, a.q, a.B, a.Lq, a.H, a.ldq, 128))) return rc;
  if ((rc = make_qkv_tmap3(&tk, a.k, a.B, a.Lk, a.H, a.ldk, BKV))) return rc;
  if ((rc = make_qkv_tmap3(&tv, a.v, a.B, a.Lk, a.H, a.ldv, BKV))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    CE_CHECK_CUDA(cudaFuncSetAttribute(attention3_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem3::total));
    attr_set = true;
  }
  dim3 grid((a.Lq + 2 * BQ - 1) / (2 * BQ), a.H, a.B);
  attention3_fwd_kernel<<<grid, ATTN3_THREADS, Smem3::total, stream>>>(tq, tk, tv, a);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
