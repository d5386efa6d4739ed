// Self-attention forward, second generation: TWO 128-query tiles per CTA share every K/V tile, P lives in TMEM.
// (Round-1 kernel with the round-2 issue path; attention6.cu -- two softmax threads per score row -- is the default now, this one
// stays selectable with CE_ATTN_V2=2 / ce_debug_attention_kernel(2) and is covered by the same tests.)
//
//   out[b, i, h*128:(h+1)*128] = softmax_j( q[b,i,h,:] . k[b,j,h,:] * scale ) @ v[b,j,h,:]          (head_dim 128)
//
// One CTA per (256 queries, head, batch); 384 threads = 3 warpgroups (setmaxnreg 80 / 208 / 208):
//   warp 0        TMA producer: Q0, Q1 once; K tiles through a 3-deep ring, V tiles through a 2-deep ring (128 keys each)
//   warp 1        MMA issuer (one elected thread, fixed order, see below): S_q = Q_q K_j^T (SS)  and  O_q += P_q V_j (TS: P read from TMEM,
//                 V consumed as an MN-major B operand straight from the row-major tile)
//   warps 4-7     softmax group 0 = query tile 0,   warps 8-11 softmax group 1 = query tile 1: ordinary online softmax over
//                 ALL key tiles (running max / sum in registers, lazy rescale of O), P written back as packed bf16 into the
//                 first 64 columns of the group's own S region (tcgen05.st) — no shared-memory round trip for P.
// TMEM (512 columns): S0|P0 [0,128)  S1|P1 [128,256)  O0 [256,384)  O1 [384,512).
// Per query tile the chain S -> softmax -> P.V -> next S is serial (P aliases S; the next S is queued right behind the
// P.V that reads P, relying on in-order execution of the MMA pipe), and the two query tiles interleave on the tensor pipe; each K/V tile is fetched ONCE per 256 queries (half the L2->SM traffic of attention.cu).
//
// Replaces F.scaled_dot_product_attention of the self-attention (transformer_chronoedit.py:97-99).  attention.cu remains
// the kernel for the two-source cross-attention and for key counts that are not worth two query tiles.
#include <cstdlib>

#include "attention.cuh"

namespace ce {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int NK = 3;  // K ring depth
constexpr int NV = 2;  // V ring depth
constexpr int ATTN2_THREADS = 384;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;
constexpr uint32_t HALF_BYTES = TILE_BYTES / 2;
constexpr float RESCALE_THRESHOLD = 8.0f;

// defaults of the A/B knobs (chosen from profiles/r2*_attention_*.log)
#define CE_ATTN_POLY_DEFAULT 1
#define CE_ATTN_SPEC_DEFAULT false
#define CE_ATTN_QUARTERS_DEFAULT false

struct Smem2 {
  static constexpr uint32_t q = 0;                         // 2 tiles
  static constexpr uint32_t k = q + 2 * TILE_BYTES;        // NK tiles
  static constexpr uint32_t v = k + NK * TILE_BYTES;       // NV tiles
  static constexpr uint32_t bars = v + NV * TILE_BYTES;
  static constexpr uint32_t total = bars + 256;
};

enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + NK, V_FULL = K_EMPTY + NK, V_EMPTY = V_FULL + NV, S_FULL = V_EMPTY + NV,
       P_FULL = S_FULL + 2 /* [part*2 + q]: part = 32-key quarter (QUARTERS) or 64-key half of the tile */, PV_DONE = P_FULL + 8,
       NUM_BARS2 = PV_DONE + 2 };

// The MMA issuer (one elected lane of warp 1) walks a FIXED order with the descriptors computed ahead of its waits and every
// group of MMAs issued as ONE asm statement: per key tile j, for query tile 0 then 1: P.V(j) part by part as the softmax
// publishes P, then S(j+1).  Why (profiles/r2c_*, r2d_*, r2f_*): (i) the issuing thread's own work between two tcgen05.mma is
// not hidden when it is longer than the MMA (64 cycles here): with inline descriptor arithmetic and the compiler's per-MMA
// serialisation loop (`if (lane == 0)` instead of elect.sync) each MMA took 95-115 cycles; (ii) the tensor pipe executes in
// issue order, so a greedy "issue whatever is ready" loop lets one tile's early P.V slip in front of the other tile's S.
// Template knobs (A/B through the environment, see launch_attention2):
//   POLY8    of every 8 exp2 pairs are evaluated on the FMA pipe (f2_exp2_poly), the rest on the MUFU;
//   SPEC     exponentials start against the PREVIOUS running maximum on the first 32 score columns while the other 96 are still
//            coming from TMEM, the tile's own maximum is folded in between the MUFU instructions, and the tile is redone the
//            classic way (max first) only when that maximum exceeds the running one by more than the lazy-rescale threshold --
//            bit-identical results, because without a rescale the classic path uses the same stale maximum;
//   QUARTERS P is published (and P.V issued) in four 32-key parts instead of two 64-key halves.
template <int POLY8, bool SPEC, bool QUARTERS>
__global__ void __launch_bounds__(ATTN2_THREADS, 1)
attention2_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem2::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS2);
  constexpr int PARTS = QUARTERS ? 4 : 2;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (a.Lk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[chronoedit_b200] attention2: dynamic shared memory not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < NUM_BARS2; ++i) mbar_init(&bars[i], (i >= P_FULL && i < P_FULL + 8) ? 128 : 1);
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
  const bool timed_blk = a.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer: K(t), V(t) in order, blocking waits
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&bars[Q_FULL], 2 * TILE_BYTES);
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
          tma_load_3d(smem + Smem2::q + qt * TILE_BYTES, &tma_q, &bars[Q_FULL], h * HD, q0 + qt * BQ, b);
          tma_load_3d(smem + Smem2::q + qt * TILE_BYTES + HALF_BYTES, &tma_q, &bars[Q_FULL], h * HD + 64, q0 + qt * BQ, b);
        }
        for (int t = 0; t < n_tiles; ++t) {
          {
            const int st = t % NK;
            mbar_wait(&bars[K_EMPTY + st], ((t / NK) & 1) ^ 1, 10 + st);
            uint8_t* ks = smem + Smem2::k + st * TILE_BYTES;
            mbar_arrive_expect_tx(&bars[K_FULL + st], TILE_BYTES);
            tma_load_3d(ks, &tma_k, &bars[K_FULL + st], h * HD, t * BKV, b);
            tma_load_3d(ks + HALF_BYTES, &tma_k, &bars[K_FULL + st], h * HD + 64, t * BKV, b);
          }
          {
            const int st = t % NV;
            mbar_wait(&bars[V_EMPTY + st], ((t / NV) & 1) ^ 1, 20 + st);
            uint8_t* vs = smem + Smem2::v + st * TILE_BYTES;
            mbar_arrive_expect_tx(&bars[V_FULL + st], TILE_BYTES);
            tma_load_3d(vs, &tma_v, &bars[V_FULL + st], h * HD, t * BKV, b);
            tma_load_3d(vs + HALF_BYTES, &tma_v, &bars[V_FULL + st], h * HD + 64, t * BKV, b);
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (one elected lane, fixed order)
      if (elect_one_sync()) {
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, 0);   // Q (K-major, smem) x K^T (K-major, smem)
        constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 1);  // P (TMEM) x V (MN-major, smem)
        mbar_wait(&bars[Q_FULL], 0, 1);
#define CE_EVT(jj, slot) \
  if (timed_blk && (jj) >= 16 && (jj) < 24) a.timing[64 + qt * 64 + ((jj)-16) * 8 + (slot)] = clock64();
        // descriptors are computed AHEAD of the waits (slack time); Q descriptors never change, K / V ones depend on the ring stage
        uint64_t dq[2][8];
#pragma unroll
        for (int qt = 0; qt < 2; ++qt)
#pragma unroll
          for (int kk = 0; kk < 8; ++kk)
            dq[qt][kk] = umma_desc_kmajor_sw128(smem_u32(smem + Smem2::q + qt * TILE_BYTES) + (kk >> 2) * HALF_BYTES) + 2 * (kk & 3);
        uint64_t dk[8], dv[8];
        auto make_dk = [&](int j) {
          const uint32_t k_addr = smem_u32(smem + Smem2::k + (j % NK) * TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) dk[kk] = umma_desc_kmajor_sw128(k_addr + (kk >> 2) * HALF_BYTES) + 2 * (kk & 3);
        };
        auto make_dv = [&](int j) {
          const uint32_t v_addr = smem_u32(smem + Smem2::v + (j % NV) * TILE_BYTES);
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) dv[kk] = umma_desc_mnmajor_sw128(v_addr + kk * 2048, HALF_BYTES);
        };
        auto issue_s = [&](int qt, int j) {   // dk must hold the descriptors of K(j)
          mbar_wait(&bars[K_FULL + j % NK], (j / NK) & 1, 30 + qt);
          tc_fence_after();
          umma_bf16_ss_x8(tmem_base + qt * 128, dq[qt], dk, IDESC_S, 0);
          umma_commit(&bars[S_FULL + qt]);
          if (qt == 1) umma_commit(&bars[K_EMPTY + j % NK]);   // both query tiles have consumed K_j
          CE_EVT(j, 5)
        };
        make_dk(0);
        issue_s(0, 0);
        issue_s(1, 0);
        for (int j = 0; j < n_tiles; ++j) {
          make_dv(j);
          if (j + 1 < n_tiles) make_dk(j + 1);
#pragma unroll
          for (int qt = 0; qt < 2; ++qt) {
            const uint32_t p_tmem = tmem_base + qt * 128;   // packed bf16: 8 columns per K=16 step
            const uint32_t d = tmem_base + 256 + qt * 128;
#pragma unroll
            for (int part = 0; part < PARTS; ++part) {
              mbar_wait(&bars[P_FULL + part * 2 + qt], j & 1, 40 + qt);   // (spinning here instead of try_wait: measured -3.5 %)
              if (part == 0 && qt == 0) mbar_wait(&bars[V_FULL + j % NV], (j / NV) & 1, 44);
              tc_fence_after();
              if (QUARTERS) umma_bf16_ts_x2(d, p_tmem + part * 16, 8, dv[2 * part], dv[2 * part + 1], IDESC_PV, (j | part) != 0);
              else umma_bf16_ts_x4(d, p_tmem + part * 32, 8, dv[4 * part], dv[4 * part + 1], dv[4 * part + 2], dv[4 * part + 3], IDESC_PV, (j | part) != 0);
              if (part == 0) { CE_EVT(j, 3) }
            }
            umma_commit(&bars[PV_DONE + qt]);
            if (qt == 1) umma_commit(&bars[V_EMPTY + j % NV]);   // both query tiles have consumed V_j
            CE_EVT(j, 4)
            if (j + 1 < n_tiles) issue_s(qt, j + 1);
          }
        }
#undef CE_EVT
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
    // ---------------------------------------------------------------- softmax groups (one per query tile)
    const int qt = (warp - 4) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + qt * 128;
    const uint32_t o_tmem = tmem_base + lane_base + 256 + qt * 128;
    const float sl2 = a.scale * 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    const bool timed = timed_blk && (threadIdx.x == 128 || threadIdx.x == 256);
    long long tacc[6] = {0, 0, 0, 0, 0, 0};
    long long tc0 = 0;
#define CE_TICK(slot)                      \
  if (timed) {                             \
    const long long _t = clock64();        \
    tacc[slot] += _t - tc0;                \
    tc0 = _t;                              \
  }
#define CE_SEVT(slot) \
  if (timed && j >= 16 && j < 24) a.timing[64 + qt * 64 + (j - 16) * 8 + (slot)] = clock64();
    if (timed) tc0 = clock64();

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = a.Lk - j * BKV;
      mbar_wait(&bars[S_FULL + qt], j & 1, 60 + qt);
      tc_fence_after();
      CE_TICK(0)
      CE_SEVT(0)
      uint32_t s[128];
      uint32_t pk[64];
      uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
      const uint64_t sl2_2 = f2_pack(sl2, sl2);
      uint64_t negm_2 = f2_pack(-m, -m);
      // packed fp32 pairs: FFMA2 for the scale-and-shift, FADD2 for the row sums; 2 MUFU (or the FMA-pipe polynomial) per pair
      auto exp_pair = [&](int i, bool ordered = false) {
        const uint64_t x2 = f2_fma(f2_pack_bits(s[2 * i], s[2 * i + 1]), sl2_2, negm_2);
        float p0, p1;
        if ((i & 7) < POLY8) {
          f2_exp2_poly(x2, p0, p1);
        } else {
          float x0, x1;
          f2_unpack(x2, x0, x1);
          p0 = ordered ? fast_exp2_ordered(x0) : fast_exp2(x0);
          p1 = ordered ? fast_exp2_ordered(x1) : fast_exp2(x1);
        }
        sum2[i & 3] = f2_add(sum2[i & 3], f2_pack(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
      };
      bool first_half_done = false;
      float mx = 0.f;
      if (SPEC && j > 0 && valid >= BKV) {
        // ---- speculative order: columns 0-31 first, the rest of the row lands while they are being exponentiated
        tmem_ld_32x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
        tmem_ld_wait();
#pragma unroll
        for (int c = 1; c < 4; ++c) tmem_ld_32x32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));
        CE_TICK(1)
        float mx4[4] = {__uint_as_float(s[0]), __uint_as_float(s[1]), __uint_as_float(s[2]), __uint_as_float(s[3])};
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          exp_pair(i, true);   // ordered: stays in front of the wait for columns 32-127
          if (i >= 2) mx4[i & 3] = fmaxf(mx4[i & 3], fmaxf(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])));
        }
        tmem_ld_wait();
#pragma unroll
        for (int i = 16; i < 32; ++i) {
          exp_pair(i);
          mx4[i & 3] = fmaxf(mx4[i & 3], fmaxf(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1])));
          const int e = 64 + 4 * (i - 16);   // columns 64-127: four per step
          mx4[(i + 1) & 3] = fmaxf(mx4[(i + 1) & 3], fmaxf(__uint_as_float(s[e]), __uint_as_float(s[e + 1])));
          mx4[(i + 2) & 3] = fmaxf(mx4[(i + 2) & 3], fmaxf(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])));
        }
        mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * sl2;
        CE_TICK(2)
        if (!__any_sync(0xffffffffu, mx > m + RESCALE_THRESHOLD)) {
          first_half_done = true;
        } else {  // some row of this warp outgrew the threshold: columns 0-63 again from TMEM (64-127 are still in registers)
#pragma unroll
          for (int c = 0; c < 2; ++c) tmem_ld_32x32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) sum2[i] = 0ull;
        }
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld_32x32(s_tmem + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32 * c]));
        tmem_ld_wait();
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
        mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
        mx *= sl2;
        CE_TICK(2)
      }
      if (!first_half_done) {
        float alpha = 1.0f;
        bool need = false;
        if (j == 0) {
          m = mx;
        } else {
          need = mx > m + RESCALE_THRESHOLD;
          if (need) {
            alpha = fast_exp2(m - mx);
            m = mx;
          }
        }
        // P.V(j-1) of this query tile completed before S(j) was even issued, so O is stable here
        if (__any_sync(0xffffffffu, need)) {
          l *= alpha;
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
        negm_2 = f2_pack(-m, -m);
        if (!QUARTERS) {
#pragma unroll
          for (int i = 0; i < 32; ++i) exp_pair(i);
        }
      }
      // P (packed bf16, 64 columns) overwrites the first half of this group's S region and is published part by part, so that
      // the tensor pipe works on P.V while the rest of the row is still being exponentiated
      tc_fence_after();
      if (QUARTERS) {
        if (!first_half_done) {
#pragma unroll
          for (int i = 0; i < 16; ++i) exp_pair(i);
        }
        tmem_st_32x16(s_tmem, &pk[0]);
        if (!first_half_done) {
#pragma unroll
          for (int i = 16; i < 20; ++i) exp_pair(i);
        }
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL + 0 + qt]);
        if (!first_half_done) {
#pragma unroll
          for (int i = 20; i < 32; ++i) exp_pair(i);
        }
        tmem_st_32x16(s_tmem + 16, &pk[16]);
#pragma unroll
        for (int i = 32; i < 36; ++i) exp_pair(i);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL + 2 + qt]);
        CE_TICK(3)
        CE_SEVT(1)
#pragma unroll
        for (int i = 36; i < 48; ++i) exp_pair(i);
        tmem_st_32x16(s_tmem + 32, &pk[32]);
#pragma unroll
        for (int i = 48; i < 52; ++i) exp_pair(i);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL + 4 + qt]);
#pragma unroll
        for (int i = 52; i < 64; ++i) exp_pair(i);
        tmem_st_32x16(s_tmem + 48, &pk[48]);
      } else {
        tmem_st_32x32(s_tmem, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
#pragma unroll
        for (int i = 32; i < 40; ++i) exp_pair(i);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL + 0 + qt]);
        CE_TICK(3)
        CE_SEVT(1)
#pragma unroll
        for (int i = 40; i < 64; ++i) exp_pair(i);
        tmem_st_32x32(s_tmem + 32, *reinterpret_cast<const uint32_t(*)[32]>(&pk[32]));
      }
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(sum2[0], sum2[1]), a0, a1);
        f2_unpack(f2_add(sum2[2], sum2[3]), b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[P_FULL + (PARTS - 1) * 2 + qt]);
      CE_TICK(4)
      CE_SEVT(2)
    }
    if (timed && qt == 0) {
      for (int i = 0; i < 5; ++i) a.timing[i] = tacc[i];
      a.timing[5] = n_tiles;
    }
#undef CE_TICK
#undef CE_SEVT

    // ---- normalise and store this query tile
    mbar_wait(&bars[PV_DONE + qt], (n_tiles - 1) & 1, 80 + qt);
    tc_fence_after();
    const float inv = 1.0f / l;
    const int row = q0 + qt * BQ + r;
    bf16* orow = a.out + ((size_t)b * a.Lq + row) * a.ldo + h * HD;
    if (a.peer_rows > 0 && row < a.Lq)   // sequence parallel: the token's owner gets the row (peer store over NVLink)
      orow = a.out_peer[row / a.peer_rows] + ((size_t)b * a.peer_rows + row % a.peer_rows) * a.ldo + a.out_col0 + h * HD;
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

int make_qkv_tmap2(CUtensorMap* m, const bf16* base, int B, int L, int H, int ld) {
  uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
  uint32_t box[3] = {64, 128, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

}  // namespace

int launch_attention2(const AttnArgs& a, cudaStream_t stream) {
  CE_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0 && a.Lk2 == 0 && a.accumulate == 0, "attention2: single source, no accumulate");
  CE_REQUIRE(a.peer_rows == 0 || (a.Lq + a.peer_rows - 1) / a.peer_rows <= 8, "attention2: at most 8 sequence-parallel peers");
  CE_REQUIRE(a.head_dim == HD, "attention2: only head_dim 128 is built");
  CE_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention2: leading dims % 8");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_tmap2(&tq, a.q, a.B, a.Lq, a.H, a.ldq))) return rc;
  if ((rc = make_qkv_tmap2(&tk, a.k, a.B, a.Lk, a.H, a.ldk))) return rc;
  if ((rc = make_qkv_tmap2(&tv, a.v, a.B, a.Lk, a.H, a.ldv))) return rc;
  // developer knobs (A/B, profiles/r2g_attention_variant_sweep.log): CE_ATTN_POLY = how many of every 8 exp2 pairs run on the FMA
  // pipe (0..2; 1 is +2.3 %, 2 is +0.7 %, 3 is 0); CE_ATTN_SPEC=1 speculative (previous-maximum) softmax order (0 +- 1 %);
  // CE_ATTN_QUARTERS=1 P in four parts (+0.7 %, nothing on top of POLY=1)
  static const int poly = [] {
    const char* e = getenv("CE_ATTN_POLY");
    const int v = e ? atoi(e) : CE_ATTN_POLY_DEFAULT;
    return v < 0 ? 0 : (v > 2 ? 2 : v);
  }();
  auto flag = [](const char* name, bool dflt) {
    const char* e = getenv(name);
    return e ? e[0] == '1' : dflt;
  };
  static const bool spec = flag("CE_ATTN_SPEC", CE_ATTN_SPEC_DEFAULT);
  static const bool quarters = flag("CE_ATTN_QUARTERS", CE_ATTN_QUARTERS_DEFAULT);
  dim3 grid((a.Lq + 2 * BQ - 1) / (2 * BQ), a.H, a.B);
#define CE_LAUNCH_ATTN2(P, S, Q)                                                                      \
  do {                                                                                                \
    CE_ENSURE_SMEM((attention2_fwd_kernel<P, S, Q>), Smem2::total);                                   \
    attention2_fwd_kernel<P, S, Q><<<grid, ATTN2_THREADS, Smem2::total, stream>>>(tq, tk, tv, a);     \
  } while (0)
#define CE_LAUNCH_ATTN2_P(S, Q)                       \
  switch (poly) {                                     \
    case 0: CE_LAUNCH_ATTN2(0, S, Q); break;          \
    case 1: CE_LAUNCH_ATTN2(1, S, Q); break;          \
    default: CE_LAUNCH_ATTN2(2, S, Q); break;         \
  }
  switch ((spec ? 2 : 0) | (quarters ? 1 : 0)) {
    case 0: CE_LAUNCH_ATTN2_P(false, false) break;
    case 1: CE_LAUNCH_ATTN2_P(false, true) break;
    case 2: CE_LAUNCH_ATTN2_P(true, false) break;
    default: CE_LAUNCH_ATTN2_P(true, true) break;
  }
#undef CE_LAUNCH_ATTN2_P
#undef CE_LAUNCH_ATTN2
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
