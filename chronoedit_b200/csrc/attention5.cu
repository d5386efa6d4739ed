// Self-attention forward with cta_group::2 MMAs: a cluster of TWO CTAs (one TPC) computes 256 queries x one head.
//   S  = [Q0;Q1] K_j^T   one M256 x N128 instruction stream issued by the leader CTA: each SM multiplies ITS 128 queries, the
//                        B operand K_j is split across the pair (each SM holds 64 of the 128 keys: 16 KB instead of 32 KB)
//   O += [P0;P1] V_j     P read from each SM's TMEM, V_j split across the pair along the head dimension (64 of 128 dims each)
// Per SM and key tile the shared memory sees 32 KB of Q + 16 KB of K + 16 KB of V operand reads and 32 KB of TMA fill =
// 96 KB, against 128 KB in attention2.cu (and 160 KB for a TMA-multicast variant of this kernel that was measured and dropped): SS MMAs at M128 x N128 x K16 already
// consume the SM's full 128 B/clk of shared-memory bandwidth, so this is what decides whether the tensor pipe can stay busy.
// Each CTA owns ONE 128-query tile and all 512 TMEM columns of its SM:
//     S0 [0,128)  S1 [128,256)   score buffers of the even / odd key tiles
//     P0 [256,320) P1 [320,384)  P (packed bf16) in its own columns: no serial S -> P -> P.V -> S chain
//     O  [384,512)               ONE accumulator
// 384 threads per CTA: warp 0 TMA producer (its halves of K/V, its Q), warp 1 MMA issuer (leader CTA only, event driven),
// warps 4-7 softmax group A = even key tiles, warps 8-11 group B = odd key tiles, one thread per query row.  Both groups
// feed the SAME accumulator, so they share the running row maximum: the thread that handles tile j reads m(j-1) from a
// per-row mailbox in shared memory (published by the other group right after its max phase), decides m(j) with the usual
// lazy threshold, publishes it, and -- in the rare case of a jump -- rescales O itself once P.V(j-1) has completed.  Only
// this short decide step is serial; the long phases (TMEM load, row max, 128 exp2, P store) of consecutive tiles overlap.
// Each thread keeps the partial row sum of its own tiles relative to the maximum it used last; the two partial sums are
// brought to the final maximum and added at the end.  Per row this is the same sequence of operations as attention2.cu
// (same threshold decisions, same bf16 P).
// Barriers that gather BOTH CTAs (operands landed, S consumed, P published) live in the leader; completion of the MMAs is
// multicast to both CTAs with tcgen05.commit.
//
// STATUS: parity-green (tests/test_gpu_ops.py::test_attention_alternative_kernels) but NOT the default: 1020-1045 TFLOP/s
// against 1123 for attention2.cu.  scripts/attn_timing.py: leader and peer show identical phase times and uncontended exp
// phases, but the two handshake loops that cross the CTA pair ("S consumed" -> S(j+2) issued -> "S ready"; "P published" ->
// P.V issued -> "P.V done") each take ~2000-2300 cycles and two S / two P buffers are all the TMEM has to hide them, so the
// period stays ~3650 cycles per two tiles wherever the waits are placed.  Selected with CE_ATTN_V2=5 or
// ce_debug_attention_kernel(5).
//
// Replaces F.scaled_dot_product_attention of the self-attention (transformer_chronoedit.py:97-99).
#include <cstdlib>

#include "attention.cuh"

namespace ce {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int NK = 4;  // K ring depth (S is issued two tiles ahead of the softmax)
constexpr int NV = 4;  // V ring depth
constexpr int ATTN5_THREADS = 384;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;
constexpr uint32_t HALF_BYTES = TILE_BYTES / 2;    // one 64-wide head-dim half of a tile (128 rows x 128 B)
constexpr uint32_t KV_BYTES = HALF_BYTES;           // this CTA's half of a K tile (64 keys x 128 dims) or V tile (128 keys x 64 dims)
constexpr uint32_t KPART_BYTES = 64 * 64 * 2;      // 64 keys x 64 dims: one head-dim half of the K half-tile
constexpr float RESCALE_THRESHOLD = 8.0f;
constexpr uint16_t BOTH_CTAS = 0x3;
constexpr uint32_t LEADER = 0;

struct Smem5 {
  static constexpr uint32_t q = 0;
  static constexpr uint32_t k = q + TILE_BYTES;
  static constexpr uint32_t v = k + NK * KV_BYTES;
  static constexpr uint32_t xchg = v + NV * KV_BYTES;   // 2 x 2 x 128 floats: [tile parity][0: running max m(j), 1: partial row sum][row]
  static constexpr uint32_t bars = xchg + 2 * 2 * 128 * 4;
  static constexpr uint32_t total = bars + 256;
};
static_assert(Smem5::total <= 227 * 1024, "attention5: shared memory budget");

enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + NK, V_FULL = K_EMPTY + NK, V_EMPTY = V_FULL + NV, S_FULL = V_EMPTY + NV, S_FREE = S_FULL + 2,
       P_FULL = S_FREE + 2, PV_DONE = P_FULL + 2, M_PUB = PV_DONE + 2, EXP_DONE = M_PUB + 2, NUM_BARS5 = EXP_DONE + 2 };
static_assert(NUM_BARS5 * 8 + 8 <= 256, "attention5: barrier block");

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(ATTN5_THREADS, 1)
attention5_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem5::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS5);
  float* xchg = reinterpret_cast<float*>(smem + Smem5::xchg);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();   // == blockIdx.x & 1
  const int q0 = blockIdx.x * BQ;             // may lie beyond Lq for the padding CTA of an odd tile count: it still feeds its peer
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (a.Lk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[chronoedit_b200] attention5: dynamic shared memory not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < NUM_BARS5; ++i) {
      uint32_t count = 1;                                                        // multicast commits, EMPTY barriers
      if (i == Q_FULL || (i >= K_FULL && i < K_FULL + NK) || (i >= V_FULL && i < V_FULL + NV)) count = 2;   // leader: one producer per CTA
      if (i >= S_FREE && i < P_FULL + 2) count = 8;                              // leader: one arrival per softmax warp of the group, both CTAs
      if (i >= M_PUB) count = 128;                                               // every thread of one softmax group (M_PUB, EXP_DONE)
      mbar_init(&bars[i], count);
    }
    fence_mbar_init();
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
  }
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers exist before any remote arrive / TMA credit
  if (warp == 1) tmem_alloc_2sm(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool leader = crank == LEADER;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 48;");
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer (event-driven)
      if (lane == 0) {
        // operands land in THIS CTA's shared memory; the byte counts are credited to the leader's barrier (cta_group::2 TMA)
        if (leader) mbar_arrive_expect_tx(&bars[Q_FULL], 2 * TILE_BYTES);
        tma_load_3d_2sm(smem + Smem5::q, &tma_q, &bars[Q_FULL], h * HD, q0, b);
        tma_load_3d_2sm(smem + Smem5::q + HALF_BYTES, &tma_q, &bars[Q_FULL], h * HD + 64, q0, b);
        if (!leader) mbar_arrive_remote(&bars[Q_FULL], LEADER);
        int k_next = 0, v_next = 0;
        uint64_t t_start = 0;
        uint32_t idle = 0;
        while (k_next < n_tiles || v_next < n_tiles) {
          bool progress = false;
          if (k_next < n_tiles) {
            const int st = k_next % NK;
            if (mbar_test_wait(&bars[K_EMPTY + st], ((k_next / NK) & 1) ^ 1)) {
              // keys [crank*64, +64) of the tile, as two 64-dim halves of 64 rows x 128 B
              uint8_t* ks = smem + Smem5::k + st * KV_BYTES;
              if (leader) mbar_arrive_expect_tx(&bars[K_FULL + st], 2 * KV_BYTES);
              tma_load_3d_2sm(ks, &tma_k, &bars[K_FULL + st], h * HD, k_next * BKV + (int)crank * 64, b);
              tma_load_3d_2sm(ks + KPART_BYTES, &tma_k, &bars[K_FULL + st], h * HD + 64, k_next * BKV + (int)crank * 64, b);
              if (!leader) mbar_arrive_remote(&bars[K_FULL + st], LEADER);
              ++k_next;
              progress = true;
            }
          }
          if (v_next < n_tiles) {
            const int st = v_next % NV;
            if (mbar_test_wait(&bars[V_EMPTY + st], ((v_next / NV) & 1) ^ 1)) {
              // head dims [crank*64, +64) of all 128 keys: one box of 128 rows x 128 B
              if (leader) mbar_arrive_expect_tx(&bars[V_FULL + st], 2 * KV_BYTES);
              tma_load_3d_2sm(smem + Smem5::v + st * KV_BYTES, &tma_v, &bars[V_FULL + st], h * HD + (int)crank * 64, v_next * BKV, b);
              if (!leader) mbar_arrive_remote(&bars[V_FULL + st], LEADER);
              ++v_next;
              progress = true;
            }
          }
          if (progress) {
            idle = 0;
          } else if ((++idle & 0xFFF) == 0) {
            if (t_start == 0) t_start = global_timer_ns();
            else if (global_timer_ns() - t_start > CE_MBAR_TIMEOUT_NS) {
              printf("[chronoedit_b200] attention5 producer stalled: block=(%d,%d,%d) k=%d v=%d\n", blockIdx.x, blockIdx.y, blockIdx.z, k_next, v_next);
              __trap();
            }
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (leader CTA only)
      if (leader && lane == 0) {
        constexpr uint32_t IDESC_S = umma_idesc_bf16(256, 128, 0);   // [Q0;Q1] (K-major, smem) x K^T (K-major, smem, 64 keys per CTA)
        constexpr uint32_t IDESC_PV = umma_idesc_bf16(256, 128, 1);  // [P0;P1] (TMEM) x V (MN-major, smem, 64 dims per CTA)
        const uint32_t q_addr = smem_u32(smem + Smem5::q);
        const uint32_t o_tm = tmem_base + 384;
        mbar_wait(&bars[Q_FULL], 0, 1);
        int s_next = 0, pv_next = 0;
        uint64_t t_start = 0;
        uint32_t idle = 0;
        while (pv_next < n_tiles) {
          bool progress = false;
          // S(j) -> buffer j&1: free once the group of that parity has pulled S(j-2) into registers (early in its step)
          if (s_next < n_tiles) {
            const int j = s_next;
            if ((j < 2 || mbar_test_wait(&bars[S_FREE + (j & 1)], ((j - 2) >> 1) & 1)) && mbar_test_wait(&bars[K_FULL + j % NK], (j / NK) & 1)) {
              tc_fence_after();
              const uint32_t k_addr = smem_u32(smem + Smem5::k + (j % NK) * KV_BYTES);
              const uint32_t d = tmem_base + (j & 1) * 128;
#pragma unroll
              for (int kk = 0; kk < HD / 16; ++kk)
                umma_bf16_ss_2sm(d, umma_desc_kmajor_sw128(q_addr + (kk >> 2) * HALF_BYTES) + 2 * (kk & 3),
                                 umma_desc_kmajor_sw128(k_addr + (kk >> 2) * KPART_BYTES) + 2 * (kk & 3), IDESC_S, kk != 0);
              umma_commit_2sm(&bars[S_FULL + (j & 1)], BOTH_CTAS);
              umma_commit_2sm(&bars[K_EMPTY + j % NK], BOTH_CTAS);
              ++s_next;
              progress = true;
            }
          }
          // O += P(k) V(k)
          {
            const int k = pv_next;
            const bool p_ok = k < s_next && mbar_test_wait(&bars[P_FULL + (k & 1)], (k >> 1) & 1);
            if (p_ok && mbar_test_wait(&bars[V_FULL + k % NV], (k / NV) & 1)) {
              tc_fence_after();
              const uint32_t v_addr = smem_u32(smem + Smem5::v + (k % NV) * KV_BYTES);
              const uint32_t p_tm = tmem_base + 256 + (k & 1) * 64;  // packed bf16: 8 columns per K=16 step
#pragma unroll
              for (int kk = 0; kk < BKV / 16; ++kk)
                umma_bf16_ts_2sm(o_tm, p_tm + kk * 8, umma_desc_mnmajor_sw128(v_addr + kk * 2048, HALF_BYTES), IDESC_PV, (k | kk) != 0);
              umma_commit_2sm(&bars[PV_DONE + (k & 1)], BOTH_CTAS);
              umma_commit_2sm(&bars[V_EMPTY + k % NV], BOTH_CTAS);
              ++pv_next;
              progress = true;
            }
          }
          if (progress) {
            idle = 0;
          } else if ((++idle & 0xFFF) == 0) {
            if (t_start == 0) t_start = global_timer_ns();
            else if (global_timer_ns() - t_start > CE_MBAR_TIMEOUT_NS) {
              printf("[chronoedit_b200] attention5 MMA stalled: block=(%d,%d,%d) s=%d pv=%d\n", blockIdx.x, blockIdx.y, blockIdx.z, s_next, pv_next);
              __trap();
            }
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ---------------------------------------------------------------- softmax groups: A = even key tiles, B = odd key tiles
    const int grp = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + grp * 128;
    const uint32_t p_tmem = tmem_base + lane_base + 256 + grp * 64;
    const uint32_t o_tmem = tmem_base + lane_base + 384;
    const float sl2 = a.scale * 1.4426950408889634f;
    float* m_box = xchg;                 // [parity][row]: m(j) of the tile with that parity
    float* l_box = xchg + 2 * 128;       // [group][row]: partial row sums at the end
    float m_mine = -INFINITY, l = 0.f;   // the maximum this thread's partial sum is relative to
    // profiling aid: phase cycles of warp 4 lane 0 in the leader CTA (slots 0-6) and in its peer (slots 8-14) of cluster 0
    const bool timed = a.timing != nullptr && blockIdx.x < 2 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 128;
    const int tslot = blockIdx.x * 8;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tc0 = 0;
#define CE_TICK(slot)                      \
  if (timed) {                             \
    const long long _t = clock64();        \
    tacc[slot] += _t - tc0;                \
    tc0 = _t;                              \
  }
    if (timed) tc0 = clock64();

    for (int j = grp; j < n_tiles; j += 2) {
      const int it = j >> 1;   // this group's iteration = phase index of its barriers
      const int valid = a.Lk - j * BKV;
      mbar_wait(&bars[S_FULL + grp], it & 1, 60 + grp);
      tc_fence_after();
      CE_TICK(0)
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld_32x32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&bars[S_FREE + grp], LEADER);  // the buffer may take S(j+2) once all 8 warps of the pair are here
      CE_TICK(1)
      if (valid < BKV) {
#pragma unroll
        for (int i = 0; i < 128; ++i) s[i] = (i < valid) ? s[i] : 0xff800000u;
      }
      float mx8[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx8[i] = __uint_as_float(s[i]);
#pragma unroll
      for (int i = 8; i < 128; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], __uint_as_float(s[i]));
      const float mx = sl2 * fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
      CE_TICK(2)
      // ---- decide m(j) from m(j-1) (the only serial step between consecutive tiles) and publish it
      float m_prev = -INFINITY;
      if (j > 0) {
        mbar_wait(&bars[M_PUB + (grp ^ 1)], ((j - 1) >> 1) & 1, 64 + grp);
        m_prev = m_box[(grp ^ 1) * 128 + r];
      }
      const bool need = j > 0 && mx > m_prev + RESCALE_THRESHOLD;
      const float m = (j == 0 || need) ? mx : m_prev;
      m_box[grp * 128 + r] = m;
      mbar_arrive(&bars[M_PUB + grp]);   // release: the store above is visible to the waiting group
      if (m != m_mine) {                 // bring this thread's partial sum to the new reference (first tile: l = 0)
        l *= fast_exp2(m_mine - m);
        m_mine = m;
      }
      CE_TICK(3)
      CE_TICK(4)
      if (__any_sync(0xffffffffu, need)) {
        // O holds the tiles up to j-1 relative to m(j-1): P.V(j-1) must have landed, P.V(j) waits for this thread's P(j).
        // The wait on the OTHER group's barrier is only safe within one phase of it: consume this group's own P.V(j-2) first
        // (commits complete in issue order, so P.V(j-3) has landed too and the parity wait below cannot alias an older phase).
        if (j >= 2) mbar_wait(&bars[PV_DONE + grp], (it - 1) & 1, 70 + grp);
        mbar_wait(&bars[PV_DONE + (grp ^ 1)], ((j - 1) >> 1) & 1, 72 + grp);
        tc_fence_after();
        const float alpha = need ? fast_exp2(m_prev - m) : 1.0f;
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
      // The exp phase is the MUFU-bound one and the two groups share the SM's MUFU: take turns.  Group g may start the
      // exponentials of tile j once the other group has finished those of tile j-1; its TMEM load, row max, decide step and P
      // store then overlap the other group's exponentials instead of both groups contending and both idling together.
      if (j > 0) mbar_wait(&bars[EXP_DONE + (grp ^ 1)], ((j - 1) >> 1) & 1, 68);
      const float neg_m = -m;
      const uint64_t sl2_2 = f2_pack(sl2, sl2), negm_2 = f2_pack(neg_m, neg_m);
      uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
      uint32_t pk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float x0, x1;
        f2_unpack(f2_fma(f2_pack_bits(s[2 * i], s[2 * i + 1]), sl2_2, negm_2), x0, x1);
        const float p0 = fast_exp2(x0), p1 = fast_exp2(x1);
        sum2[i & 3] = f2_add(sum2[i & 3], f2_pack(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
      }
      mbar_arrive(&bars[EXP_DONE + grp]);
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(sum2[0], sum2[1]), a0, a1);
        f2_unpack(f2_add(sum2[2], sum2[3]), b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
      // this group's P buffer was last read by P.V(j-2): only the STORE needs it, so the cross-CTA round trip of that P.V
      // (peer's "P published" -> leader issues -> commit multicast back) hides behind this tile's load / max / decide / exp
      if (j >= 2) mbar_wait(&bars[PV_DONE + grp], (it - 1) & 1, 70 + grp);
      CE_TICK(5)
      tc_fence_after();
      tmem_st_32x32(p_tmem, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
      tmem_st_32x32(p_tmem + 32, *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_remote(&bars[P_FULL + grp], LEADER);
      CE_TICK(6)
    }
    if (timed) {
      for (int i = 0; i < 7; ++i) a.timing[tslot + i] = tacc[i];
      a.timing[tslot + 7] = (n_tiles + 1) / 2;
    }

    // ---- combine the two partial row sums at the final maximum; group A normalises and stores the row
    const int last = n_tiles - 1;
    mbar_wait(&bars[M_PUB + (last & 1)], (last >> 1) & 1, 66 + grp);
    const float m_final = m_box[(last & 1) * 128 + r];
    l_box[grp * 128 + r] = l * fast_exp2(m_mine - m_final);   // a group that saw no tile: l = 0, m_mine = -inf -> 0 * 0
    named_bar_sync(1 + quad, 64);
    if (grp == 0) {
      const float inv = 1.0f / (l_box[r] + l_box[128 + r]);
      mbar_wait(&bars[PV_DONE + (last & 1)], (last >> 1) & 1, 80);  // commits complete in issue order
      tc_fence_after();
      const int row = q0 + r;
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
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA may exit (or free TMEM) while the pair still reads its shared memory / signals its barriers
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

int make_qkv_tmap5(CUtensorMap* m, const bf16* base, int B, int L, int H, int ld, uint32_t box_rows) {
  uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
  uint32_t box[3] = {64, box_rows, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

}  // namespace

int launch_attention5(const AttnArgs& a, cudaStream_t stream) {
  CE_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0 && a.Lk2 == 0 && a.accumulate == 0, "attention5: single source, no accumulate");
  CE_REQUIRE(a.head_dim == HD, "attention5: only head_dim 128 is built");
  CE_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention5: leading dims % 8");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_tmap5(&tq, a.q, a.B, a.Lq, a.H, a.ldq, 128))) return rc;
  if ((rc = make_qkv_tmap5(&tk, a.k, a.B, a.Lk, a.H, a.ldk, 64))) return rc;
  if ((rc = make_qkv_tmap5(&tv, a.v, a.B, a.Lk, a.H, a.ldv, 128))) return rc;
  CE_ENSURE_SMEM(attention5_fwd_kernel, Smem5::total);
  const int q_tiles = (a.Lq + BQ - 1) / BQ;
  dim3 grid(2 * ((q_tiles + 1) / 2), a.H, a.B);   // whole clusters: an odd tile count gets one padding CTA
  attention5_fwd_kernel<<<grid, ATTN5_THREADS, Smem5::total, stream>>>(tq, tk, tv, a);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
