// Self-attention forward, third generation: attention2.cu's CTA (two 128-query tiles share every K/V tile, P in TMEM, one elected
// issuer walking a fixed order) with TWO softmax threads per score row.
//
// Why (profiles/r2f_attention_elect_batched_issue.log): in attention2.cu the period of a key tile is exactly the serial chain of ONE
// query tile -- S ready -> softmax (ld 110 + max 300 + exponentials 1580 cycles) -> last P.V part (256) -> next S (512) -> S ready -- and
// the softmax of a 128-wide row by one thread is latency bound: its single warp per scheduler keeps the MUFU only ~57 % busy.  Here a
// row is split between two threads (64 keys each, warps on the same scheduler): TMEM load, row maximum and publish latencies halve and
// two warps feed the MUFU.  The halves exchange their maxima through shared memory (one named barrier per key tile, which also orders
// "half 0 has read S columns 32-63" before "half 1 overwrites them with P") and add their row sums at the end.
//
// 640 threads = 20 warps, five per scheduler: launched at 96 registers; setmaxnreg 64 (helpers) / 104 (softmax).  The register file
// is per scheduler (16 K registers) and setmaxnreg.inc can only take what the warps of the same CTA on that scheduler released:
// one helper 96 -> 64 frees 32 x 32, four softmax warps 96 -> 104 take 4 x 8 x 32 -- exactly that (anything larger deadlocks in
// setmaxnreg.inc; 18 warps at 112 registers do not fit either: 5 warps x 112 x 32 > 16 K on the scheduler that gets five).
//   warp 0        TMA producer (Q0, Q1 once; K ring 3 deep, V ring 2 deep)
//   warp 1        MMA issuer: per key tile j, for query tile 0 then 1: P.V(j) part by part in the order the halves publish, then S(j+1)
//   warps 4-19    softmax: warp = 4 + tile*8 + half*4 + quad; thread (quad, lane) owns row quad*32+lane, keys [half*64, half*64+64)
// TMEM (512 columns): S0|P0 [0,128)  S1|P1 [128,256)  O0 [256,384)  O1 [384,512); P (packed bf16) occupies S columns [0,64).
//
// Replaces F.scaled_dot_product_attention of the self-attention (transformer_chronoedit.py:97-99).
#include <cstdlib>

#include "attention.cuh"

namespace ce {

namespace {

constexpr int HD = 128;
constexpr int BQ = 128;
constexpr int BKV = 128;
constexpr int NK = 3;
constexpr int NV = 2;
constexpr int ATTN6_THREADS = 640;
constexpr uint32_t TILE_BYTES = 128 * 128 * 2;
constexpr uint32_t HALF_BYTES = TILE_BYTES / 2;
constexpr float RESCALE_THRESHOLD = 8.0f;

struct Smem6 {
  static constexpr uint32_t q = 0;
  static constexpr uint32_t k = q + 2 * TILE_BYTES;
  static constexpr uint32_t v = k + NK * TILE_BYTES;
  static constexpr uint32_t xchg = v + NV * TILE_BYTES;    // float [tile 2][half 2][row 128] (2 KB: all that is left of the 227 KB)
  static constexpr uint32_t bars = xchg + 2 * 2 * 128 * 4;
  static constexpr uint32_t total = bars + 256;
};

enum { Q_FULL = 0, K_FULL = 1, K_EMPTY = K_FULL + NK, V_FULL = K_EMPTY + NK, V_EMPTY = V_FULL + NV, S_FULL = V_EMPTY + NV,
       P_FULL = S_FULL + 2 /* [part*2 + tile], part = half*2 + (first | second 32 keys of the half) */, PV_DONE = P_FULL + 8,
       NUM_BARS6 = PV_DONE + 2 };

// POLY8 of every 8 exp2 pairs run on the FMA pipe (f2_exp2_poly).  SPLIT: 0 = each half publishes its 64 keys of P at once; 16 / 24 = in
// two parts, the first after SPLIT exp2 pairs (32 + 32 or 48 + 16 keys): the shorter the last part, the less P.V sits in the serial chain.
template <int POLY8, int SPLIT, bool TIMED>
__global__ void __launch_bounds__(ATTN6_THREADS, 1)
attention6_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                      const __grid_constant__ CUtensorMap tma_v, AttnArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Smem6::bars);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + NUM_BARS6);
  float* xchg = reinterpret_cast<float*>(smem + Smem6::xchg);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 2 * BQ;
  const int h = blockIdx.y;
  const int b = blockIdx.z;
  const int n_tiles = (a.Lk + BKV - 1) / BKV;

  if (threadIdx.x == 0) {
    if ((smem_u32(smem) & 1023u) != 0) {
      printf("[chronoedit_b200] attention6: dynamic shared memory not 1024-byte aligned\n");
      __trap();
    }
    for (int i = 0; i < NUM_BARS6; ++i) mbar_init(&bars[i], (i >= P_FULL && i < P_FULL + 8) ? 128 : 1);
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
  // TIMED: event log of block 0, key tiles 16..23 (SM clock) -> timing[64 + tile*128 + half*64 + (j-16)*8 + slot]:
  //   softmax thread (quad 0, lane 0) of each half: 0 S seen, 1 S in registers, 2 row maximum exchanged, 3 first P part published,
  //   4 last P part published;   issuer (into half 0's slots): 5 first P.V part issued, 6 last P.V part issued, 7 S(j+1) issued
  const bool timed_blk = TIMED && a.timing != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0;
#define CE_EVT6(cond, qt_, hf_, jj, slot) \
  if (TIMED && (cond) && (jj) >= 16 && (jj) < 24) a.timing[64 + (qt_) * 128 + (hf_) * 64 + ((jj)-16) * 8 + (slot)] = clock64();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0) {
      // ---------------------------------------------------------------- TMA producer
      if (elect_one_sync()) {
        mbar_arrive_expect_tx(&bars[Q_FULL], 2 * TILE_BYTES);
#pragma unroll
        for (int qt = 0; qt < 2; ++qt) {
          tma_load_3d(smem + Smem6::q + qt * TILE_BYTES, &tma_q, &bars[Q_FULL], h * HD, q0 + qt * BQ, b);
          tma_load_3d(smem + Smem6::q + qt * TILE_BYTES + HALF_BYTES, &tma_q, &bars[Q_FULL], h * HD + 64, q0 + qt * BQ, b);
        }
        for (int t = 0; t < n_tiles; ++t) {
          const CUtensorMap* mk = &tma_k;
          const CUtensorMap* mv = &tma_v;
          const int key0 = t * BKV;
          {
            const int st = t % NK;
            mbar_wait(&bars[K_EMPTY + st], ((t / NK) & 1) ^ 1, 10 + st);
            uint8_t* ks = smem + Smem6::k + st * TILE_BYTES;
            mbar_arrive_expect_tx(&bars[K_FULL + st], TILE_BYTES);
            tma_load_3d(ks, mk, &bars[K_FULL + st], h * HD, key0, b);
            tma_load_3d(ks + HALF_BYTES, mk, &bars[K_FULL + st], h * HD + 64, key0, b);
          }
          {
            const int st = t % NV;
            mbar_wait(&bars[V_EMPTY + st], ((t / NV) & 1) ^ 1, 20 + st);
            uint8_t* vs = smem + Smem6::v + st * TILE_BYTES;
            mbar_arrive_expect_tx(&bars[V_FULL + st], TILE_BYTES);
            tma_load_3d(vs, mv, &bars[V_FULL + st], h * HD, key0, b);
            tma_load_3d(vs + HALF_BYTES, mv, &bars[V_FULL + st], h * HD + 64, key0, b);
          }
        }
      }
    } else if (warp == 1) {
      // ---------------------------------------------------------------- MMA issuer (one elected lane, fixed order)
      if (elect_one_sync()) {
        constexpr uint32_t IDESC_S = umma_idesc_bf16(128, 128, 0);
        constexpr uint32_t IDESC_PV = umma_idesc_bf16(128, 128, 1);
        mbar_wait(&bars[Q_FULL], 0, 1);
        // Descriptors are built from one base per operand right before each batch of MMAs (a handful of integer adds in front of the
        // asm statement, none between the MMAs): the issuer has to live in 64 registers (see the setmaxnreg note above), and a spilled
        // descriptor costs an L2 round trip (local memory is not kept in the 4 KB of L1 this kernel leaves) right on the critical chain
        // -- measured: 400-500 cycles per key tile (profiles/r2s_attention6_event_log.log).
        const uint32_t hi_k = uint32_t(umma_desc_kmajor_sw128(0) >> 32), hi_v = uint32_t(umma_desc_mnmajor_sw128(0, HALF_BYTES) >> 32);
        auto desc = [](uint32_t lo, uint32_t hi) { return (uint64_t(hi) << 32) | lo; };
        const uint32_t q_lo0 = uint32_t(umma_desc_kmajor_sw128(smem_u32(smem + Smem6::q)));
        const uint32_t k_lo0 = uint32_t(umma_desc_kmajor_sw128(smem_u32(smem + Smem6::k)));
        const uint32_t v_lo0 = uint32_t(umma_desc_mnmajor_sw128(smem_u32(smem + Smem6::v), HALF_BYTES));
        // (shared-memory window < 256 KB: adding byte offsets >> 4 to the 14-bit address field never carries out of it)
        auto issue_s = [&](int qt, int j) {
          mbar_wait(&bars[K_FULL + j % NK], (j / NK) & 1, 30 + qt);
          tc_fence_after();
          uint32_t qb = q_lo0 + qt * (TILE_BYTES >> 4), kb = k_lo0 + (j % NK) * (TILE_BYTES >> 4);
          asm volatile("" : "+r"(qb), "+r"(kb));   // opaque: no descriptor is carried (= spilled) from one batch of MMAs to the next
          uint64_t da[8], db[8];
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const uint32_t off = (kk >> 2) * (HALF_BYTES >> 4) + 2 * (kk & 3);
            da[kk] = desc(qb + off, hi_k);
            db[kk] = desc(kb + off, hi_k);
          }
          umma_bf16_ss_x8(tmem_base + qt * 128, da, db, IDESC_S, 0);
          umma_commit(&bars[S_FULL + qt]);
          if (qt == 1) umma_commit(&bars[K_EMPTY + j % NK]);
        };
        issue_s(0, 0);
        issue_s(1, 0);
        for (int j = 0; j < n_tiles; ++j) {
          const uint32_t vb0 = v_lo0 + (j % NV) * (TILE_BYTES >> 4);   // V(j): K-step kk starts 2048 bytes (16 keys x 128 B) further on
          const bool fresh = j == 0;                                   // the first P.V overwrites O
#pragma unroll
          for (int qt = 0; qt < 2; ++qt) {
            const uint32_t p_tmem = tmem_base + qt * 128;   // packed bf16: 8 columns per K=16 step
            const uint32_t d = tmem_base + 256 + qt * 128;
            if (SPLIT != 0) {
              // the two halves run side by side: their first parts arrive first, then their last parts
              constexpr int F = SPLIT / 8;   // K=16 steps in the first part of a half (2 or 3 of its 4)
#pragma unroll
              for (int o = 0; o < 4; ++o) {
                const int hf = o & 1, last = o >> 1, part = hf * 2 + last;
                const int ks0 = hf * 4 + (last ? F : 0), nks = last ? 4 - F : F;
                mbar_wait(&bars[P_FULL + part * 2 + qt], j & 1, 40 + qt);
                if (o == 0 && qt == 0) mbar_wait(&bars[V_FULL + j % NV], (j / NV) & 1, 44);
                tc_fence_after();
                uint32_t vb = vb0;
                asm volatile("" : "+r"(vb));
                if (nks >= 2) umma_bf16_ts_x2(d, p_tmem + ks0 * 8, 8, desc(vb + ks0 * 128, hi_v), desc(vb + (ks0 + 1) * 128, hi_v), IDESC_PV, !(fresh && o == 0));
                if (nks == 1) umma_bf16_ts(d, p_tmem + ks0 * 8, desc(vb + ks0 * 128, hi_v), IDESC_PV, 1);
                if (nks == 3) umma_bf16_ts(d, p_tmem + (ks0 + 2) * 8, desc(vb + (ks0 + 2) * 128, hi_v), IDESC_PV, 1);
                if (o == 0) { CE_EVT6(timed_blk, qt, 0, j, 5) }
              }
            } else {
#pragma unroll
              for (int hf = 0; hf < 2; ++hf) {
                mbar_wait(&bars[P_FULL + (hf * 2) * 2 + qt], j & 1, 40 + qt);
                if (hf == 0 && qt == 0) mbar_wait(&bars[V_FULL + j % NV], (j / NV) & 1, 44);
                tc_fence_after();
                uint32_t vb = vb0;
                asm volatile("" : "+r"(vb));
                umma_bf16_ts_x4(d, p_tmem + hf * 32, 8, desc(vb + (4 * hf) * 128, hi_v), desc(vb + (4 * hf + 1) * 128, hi_v), desc(vb + (4 * hf + 2) * 128, hi_v),
                                desc(vb + (4 * hf + 3) * 128, hi_v), IDESC_PV, !(fresh && hf == 0));
              }
            }
            umma_commit(&bars[PV_DONE + qt]);
            if (qt == 1) umma_commit(&bars[V_EMPTY + j % NV]);
            CE_EVT6(timed_blk, qt, 0, j, 6)
            if (j + 1 < n_tiles) issue_s(qt, j + 1);
            CE_EVT6(timed_blk, qt, 0, j, 7)
          }
        }
      }
    }
  } else {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    // ---------------------------------------------------------------- softmax: two threads per score row
    const int sw = warp - 4;
    const int qt = sw >> 3;
    const int hf = (sw >> 2) & 1;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    const uint32_t s_tmem = tmem_base + lane_base + qt * 128 + hf * 64;         // this thread's 64 score columns
    const uint32_t p_tmem = tmem_base + lane_base + qt * 128 + hf * 32;         // ... and where its packed P goes
    const uint32_t o_tmem = tmem_base + lane_base + 256 + qt * 128 + hf * 64;   // its half of the O columns (rescale, epilogue)
    float* xs = xchg + qt * 256;                                                // [half][row]
    const float sl2 = a.scale * 1.4426950408889634f;
    float m = -INFINITY, l = 0.f;
    const bool timed = timed_blk && quad == 0 && lane == 0;

    for (int j = 0; j < n_tiles; ++j) {
      const int valid = a.Lk - j * BKV - hf * 64;   // keys of this half that exist (may be <= 0 in the last tile)
      mbar_wait(&bars[S_FULL + qt], j & 1, 60 + qt);
      tc_fence_after();
      CE_EVT6(timed, qt, hf, j, 0)
      uint32_t s[64];
      uint32_t pk[32];
      tmem_ld_32x32(s_tmem, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
      tmem_ld_32x32(s_tmem + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
      tmem_ld_wait();
      CE_EVT6(timed, qt, hf, j, 1)
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) s[i] = (i < valid) ? s[i] : 0xff800000u;
      }
      float mx4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) mx4[i] = __uint_as_float(s[i]);
#pragma unroll
      for (int i = 4; i < 64; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(s[i]));
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      // exchange with the thread that holds the other 64 keys of this row; the barrier also orders "half 0 has its S columns in
      // registers" before half 1 overwrites S columns 32-63 with its P.  One slot per thread is enough: the partner reads it before
      // it arrives on P_FULL(j), and this thread's next write follows S_FULL(j+1), which the issuer commits after P_FULL(j).
      xs[hf * 128 + r] = mx;
      named_bar_sync(1 + qt * 4 + quad, 64);   // just the two warps that share these 32 rows
      mx = fmaxf(mx, xs[(hf ^ 1) * 128 + r]) * sl2;
      CE_EVT6(timed, qt, hf, j, 2)
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
      // P.V(j-1) of this query tile completed before S(j) was issued, so O is stable here; each half rescales its 64 columns
      if (__any_sync(0xffffffffu, need)) {
        l *= alpha;
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
      const uint64_t sl2_2 = f2_pack(sl2, sl2);
      const uint64_t negm_2 = f2_pack(-m, -m);
      uint64_t sum2[4] = {0ull, 0ull, 0ull, 0ull};
      auto exp_pair = [&](int i) {
        const uint64_t x2 = f2_fma(f2_pack_bits(s[2 * i], s[2 * i + 1]), sl2_2, negm_2);
        float p0, p1;
        if ((i & 7) < POLY8) {
          f2_exp2_poly(x2, p0, p1);
        } else {
          float x0, x1;
          f2_unpack(x2, x0, x1);
          p0 = fast_exp2(x0);
          p1 = fast_exp2(x1);
        }
        sum2[i & 3] = f2_add(sum2[i & 3], f2_pack(p0, p1));
        pk[i] = pack_bf16x2(p0, p1);
      };
      tc_fence_after();
      if (SPLIT != 0) {
        constexpr int COVER = SPLIT == 16 ? 6 : 4;   // exp2 pairs issued between the P store and the wait for it
#pragma unroll
        for (int i = 0; i < SPLIT; ++i) exp_pair(i);
        tmem_st_32x16(p_tmem, &pk[0]);
        if (SPLIT == 24) tmem_st_32x8(p_tmem + 16, &pk[16]);
#pragma unroll
        for (int i = SPLIT; i < SPLIT + COVER; ++i) exp_pair(i);
        tmem_st_wait();
        tc_fence_before();
        mbar_arrive(&bars[P_FULL + (hf * 2) * 2 + qt]);
        CE_EVT6(timed, qt, hf, j, 3)
#pragma unroll
        for (int i = SPLIT + COVER; i < 32; ++i) exp_pair(i);
        if (SPLIT == 16) tmem_st_32x16(p_tmem + 16, &pk[16]);
        else tmem_st_32x8(p_tmem + 24, &pk[24]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) exp_pair(i);
        tmem_st_32x32(p_tmem, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
      }
      {
        float a0, a1, b0, b1;
        f2_unpack(f2_add(sum2[0], sum2[1]), a0, a1);
        f2_unpack(f2_add(sum2[2], sum2[3]), b0, b1);
        l += (a0 + a1) + (b0 + b1);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&bars[P_FULL + (hf * 2 + (SPLIT != 0 ? 1 : 0)) * 2 + qt]);
      CE_EVT6(timed, qt, hf, j, 4)
    }

    // ---- row sum of both halves, normalise, store this thread's 64 output columns
    mbar_wait(&bars[PV_DONE + qt], (n_tiles - 1) & 1, 80 + qt);   // (also: the partner has read the last maximum from the slot)
    tc_fence_after();
    xs[hf * 128 + r] = l;
    named_bar_sync(1 + qt * 4 + quad, 64);
    const float inv = 1.0f / (xs[r] + xs[128 + r]);   // fixed order (half 0 + half 1): both threads of a row use the same sum
    const int row = q0 + qt * BQ + r;
    bf16* orow = a.out + ((size_t)b * a.Lq + row) * a.ldo + h * HD + hf * 64;
    if (a.peer_rows > 0 && row < a.Lq)   // sequence parallel: the token's owner gets the row (peer store over NVLink)
      orow = a.out_peer[row / a.peer_rows] + ((size_t)b * a.peer_rows + row % a.peer_rows) * a.ldo + a.out_col0 + h * HD + hf * 64;
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

#undef CE_EVT6
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

int make_qkv_tmap6(CUtensorMap* m, const bf16* base, int B, int L, int H, int ld) {
  uint64_t dims[3] = {(uint64_t)H * HD, (uint64_t)L, (uint64_t)B};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)L * ld * 2};
  uint32_t box[3] = {64, 128, 1};
  return make_tmap_bf16(m, base, 3, dims, strides, box);
}

}  // namespace

int launch_attention6(const AttnArgs& a, cudaStream_t stream) {
  CE_REQUIRE(a.B > 0 && a.H > 0 && a.Lq > 0 && a.Lk > 0 && a.Lk2 == 0 && a.accumulate == 0, "attention6: single source, no accumulate");
  CE_REQUIRE(a.peer_rows == 0 || (a.Lq + a.peer_rows - 1) / a.peer_rows <= 8, "attention6: at most 8 sequence-parallel peers");
  CE_REQUIRE(a.head_dim == HD, "attention6: only head_dim 128 is built");
  CE_REQUIRE(a.ldq % 8 == 0 && a.ldk % 8 == 0 && a.ldv % 8 == 0 && a.ldo % 8 == 0, "attention6: leading dims % 8");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_qkv_tmap6(&tq, a.q, a.B, a.Lq, a.H, a.ldq))) return rc;
  if ((rc = make_qkv_tmap6(&tk, a.k, a.B, a.Lk, a.H, a.ldk))) return rc;
  if ((rc = make_qkv_tmap6(&tv, a.v, a.B, a.Lk, a.H, a.ldv))) return rc;
  // developer knobs: CE_ATTN6_POLY (0..3 of every 8 exp2 pairs on the FMA pipe), CE_ATTN6_SPLIT (0, 16, 24: see the template comment)
  static const int poly = [] {
    const char* e = getenv("CE_ATTN6_POLY");
    const int v = e ? atoi(e) : 1;
    return v < 0 ? 0 : (v > 3 ? 3 : v);
  }();
  static const int split = [] {
    const char* e = getenv("CE_ATTN6_SPLIT");
    const int v = e ? atoi(e) : 16;
    return v == 0 ? 0 : (v == 24 ? 24 : 16);
  }();
  dim3 grid((a.Lq + 2 * BQ - 1) / (2 * BQ), a.H, a.B);
#define CE_LAUNCH_ATTN6(P, SP, T)                                                                   \
  do {                                                                                                \
    CE_ENSURE_SMEM((attention6_fwd_kernel<P, SP, T>), Smem6::total);                                  \
    attention6_fwd_kernel<P, SP, T><<<grid, ATTN6_THREADS, Smem6::total, stream>>>(tq, tk, tv, a);     \
  } while (0)
#define CE_LAUNCH_ATTN6_S(SP)                         \
  switch (poly) {                                     \
    case 0: CE_LAUNCH_ATTN6(0, SP, false); break;     \
    case 1: CE_LAUNCH_ATTN6(1, SP, false); break;     \
    case 2: CE_LAUNCH_ATTN6(2, SP, false); break;     \
    default: CE_LAUNCH_ATTN6(3, SP, false); break;    \
  }
  if (a.timing) {
    if (split == 24) CE_LAUNCH_ATTN6(1, 24, true); else CE_LAUNCH_ATTN6(1, 16, true);   // event log: POLY 1 only
  } else if (split == 0) {
    CE_LAUNCH_ATTN6_S(0)
  } else if (split == 24) {
    CE_LAUNCH_ATTN6_S(24)
  } else {
    CE_LAUNCH_ATTN6_S(16)
  }
#undef CE_LAUNCH_ATTN6_S
#undef CE_LAUNCH_ATTN6
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
