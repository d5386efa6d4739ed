// HBM-bound row kernels of the DiT path: fp32 LayerNorm + adaLN modulate, RMSNorm-across-heads + 3D RoPE,
// patchify / unpatchify, and the tiny-M Linear layers of the time embedder.  All are coalesced 16-byte
// accesses with one CTA per token row (rows >> SM count), statistics in fp32 via warp shuffles.
#include "elementwise.cuh"

namespace ce {

namespace {

constexpr int ROW_THREADS = 256;
constexpr int MAX_D = 8192;                // 64 lanes x 16 vectors x 8 bf16

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  return make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}

// ---------------------------------------------------------------------------------------------
// One WARP PAIR per token row (4 rows per 256-thread CTA): each lane keeps HV 16-byte vectors of the row in registers,
// all of them loaded before anything consumes them; statistics = warp shuffles + one 64-thread named barrier.
// ~100 registers/thread -> 16+ resident warps per SM, so one row's arithmetic overlaps other rows' loads.
__device__ __forceinline__ float2 pair_sum(float a, float b, float2* part) {
  a = warp_sum(a);
  b = warp_sum(b);
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) part[warp] = make_float2(a, b);
  named_bar_sync(1 + (warp >> 1), 64);
  const float2 o = part[warp ^ 1];
  return make_float2(a + o.x, b + o.y);
}

// ---------------------------------------------------------------------------------------------
// Row statistics that arrive as per-(row, N tile) partials from the epilogue of the GEMM that produced the row (GemmArgs::stats_out):
// each warp merges them with shuffles -- Chan's pairwise update for (mean, M2), always (lower lane, higher lane) so that every lane
// ends with bit-identical values -- and the row kernel becomes a pure streaming pass: no reduction over the data, no barrier, loads
// and stores of different vectors independent of each other.
struct RowStat {
  float n, mean, m2;
};
__device__ __forceinline__ RowStat chan_merge(const RowStat& a, const RowStat& b) {
  const float tot = a.n + b.n;
  if (tot == 0.f) return a;
  const float delta = b.mean - a.mean, w = b.n / tot;
  RowStat r;
  r.n = tot;
  r.mean = fmaf(delta, w, a.mean);
  r.m2 = a.m2 + b.m2 + delta * delta * a.n * w;
  return r;
}
__device__ __forceinline__ RowStat warp_merge_stats(const float2* __restrict__ partial, int tiles, int tile_n, int D) {
  const int lane = threadIdx.x & 31;
  RowStat s{0.f, 0.f, 0.f};
  if (lane < tiles) {
    const float2 p = partial[lane];
    const int n = D - lane * tile_n;
    s.n = (float)(n < tile_n ? n : tile_n);
    s.mean = p.x;
    s.m2 = p.y;
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    RowStat t;
    t.n = __shfl_xor_sync(0xffffffffu, s.n, o);
    t.mean = __shfl_xor_sync(0xffffffffu, s.mean, o);
    t.m2 = __shfl_xor_sync(0xffffffffu, s.m2, o);
    s = (lane & o) ? chan_merge(t, s) : chan_merge(s, t);
  }
  return s;
}

template <int HV>
__global__ void __launch_bounds__(ROW_THREADS, 2)
layernorm_stats_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy, int rows, int D, float eps,
                       const float* __restrict__ scale, const float* __restrict__ shift, int mod_stride, int rows_per_batch,
                       const float* __restrict__ weight, const float* __restrict__ bias, int scale_is_1p,
                       const float2* __restrict__ stats, int stats_ld, int tiles, int tile_n) {
  const int lane64 = threadIdx.x & 63;
  const int row = blockIdx.x * (ROW_THREADS / 64) + (threadIdx.x >> 6);
  if (row >= rows) return;
  const RowStat st = warp_merge_stats(stats + (size_t)row * stats_ld, tiles, tile_n, D);
  const float mean = st.mean;
  const float rstd = rsqrtf(st.m2 / (float)D + eps);
  const float nmr = -mean * rstd;
  const int nvec = D >> 3;
  const int b = row / rows_per_batch;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * ldx);
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * ldy);
  const float* pa = scale ? scale + (size_t)b * mod_stride : weight;
  const float* pb = scale ? shift + (size_t)b * mod_stride : bias;
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    const int idx = lane64 + i * 64;
    if (idx < nvec) {
      float f[8], o[8];
      unpack8(xr[idx], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(f[j], rstd, nmr);
      if (pa) {
        const float4* qa = reinterpret_cast<const float4*>(pa + idx * 8);
        const float4* qb = reinterpret_cast<const float4*>(pb + idx * 8);
        const float4 a0 = qa[0], a1 = qa[1], b0 = qb[0], b1 = qb[1];
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        if (scale && !scale_is_1p) {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], 1.0f + av[j], bv[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], av[j], bv[j]);
        }
      }
      yr[idx] = pack8(o);
    }
  }
}

// RMSNorm across heads (+ RoPE) from epilogue partial sums of squares; `nmat` matrices side by side in one row (q | k of the fused
// QKV output): blockIdx.y selects the matrix (column offset y * D, its own weight and its own tile range)
template <int HV>
__global__ void __launch_bounds__(ROW_THREADS, 2)
rmsnorm_rope_stats_kernel(bf16* __restrict__ x, int ldx, int rows, int D, float eps, const bf16* __restrict__ weight0,
                          const bf16* __restrict__ weight1, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int L,
                          int head_dim, const float2* __restrict__ stats, int stats_ld, int tiles, SpScatter sp) {
  const int lane64 = threadIdx.x & 63;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (ROW_THREADS / 64) + (threadIdx.x >> 6);
  if (row >= rows) return;
  const int mat = blockIdx.y;
  float ss = lane < tiles ? stats[(size_t)row * stats_ld + mat * tiles + lane].x : 0.f;
  ss = warp_sum(ss);   // butterfly of commutative adds: identical in every lane
  const float rstd = rsqrtf(ss / (float)D + eps);
  const int nvec = D >> 3;
  uint4* xr = reinterpret_cast<uint4*>(x + (size_t)row * ldx + (size_t)mat * D);
  const uint4* wr = reinterpret_cast<const uint4*>(mat ? weight1 : weight0);
  const int tok = sp.world ? sp.tok0 + row % sp.rows_per_batch : (rope_cos ? row % L : 0);   // global token (RoPE position)
  const int half = head_dim >> 1;
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    const int idx = lane64 + i * 64;
    if (idx < nvec) {
      float f[8], w[8], o[8];
      unpack8(xr[idx], f);
      unpack8(wr[idx], w);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = bf16_round(bf16_round(f[j] * rstd) * w[j]);
      const int c = idx * 8;
      if (rope_cos) {
        const int pair0 = (c % head_dim) >> 1;
        const float4 cs = *reinterpret_cast<const float4*>(rope_cos + (size_t)tok * half + pair0);
        const float4 sn = *reinterpret_cast<const float4*>(rope_sin + (size_t)tok * half + pair0);
        const float cv[4] = {cs.x, cs.y, cs.z, cs.w};
        const float sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float re = o[2 * p], im = o[2 * p + 1];
          o[2 * p] = re * cv[p] - im * sv[p];
          o[2 * p + 1] = re * sv[p] + im * cv[p];
        }
      }
      if (sp.world) {   // to the rank that owns this head (peer store)
        const int head = c / head_dim, dr = head / sp.heads_per_rank, bi = row / sp.rows_per_batch;
        bf16* d = sp.dst[mat][dr] + ((size_t)bi * sp.L_total + tok) * ((size_t)sp.heads_per_rank * head_dim) +
                  (size_t)(head % sp.heads_per_rank) * head_dim + c % head_dim;
        *reinterpret_cast<uint4*>(d) = pack8(o);
      } else {
        xr[idx] = pack8(o);
      }
    }
  }
}

template <int HV>
__global__ void __launch_bounds__(ROW_THREADS, 2)
layernorm_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy, int rows, int D, float eps,
                 const float* __restrict__ scale, const float* __restrict__ shift, int mod_stride, int rows_per_batch,
                 const float* __restrict__ weight, const float* __restrict__ bias, int scale_is_1p) {
  __shared__ float2 part[ROW_THREADS / 32];
  const int lane64 = threadIdx.x & 63;
  int row = blockIdx.x * (ROW_THREADS / 64) + (threadIdx.x >> 6);
  const bool live = row < rows;
  if (!live) row = rows - 1;  // keep the pair barrier balanced; results are not stored
  const int nvec = D >> 3;
  const uint4* xr = reinterpret_cast<const uint4*>(x + (size_t)row * ldx);
  uint4 v[HV];
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    const int idx = lane64 + i * 64;
    v[i] = idx < nvec ? xr[idx] : make_uint4(0u, 0u, 0u, 0u);
  }
  // single statistics pass on pivot-shifted data d = x - x[row, 0]:  mean = pivot + E[d],  var = E[d^2] - E[d]^2
  const float pivot = __bfloat162float(x[(size_t)row * ldx]);
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    if (lane64 + i * 64 < nvec) {
      float f[8];
      unpack8(v[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[j] - pivot;
        s += d;
        ss = fmaf(d, d, ss);
      }
    }
  }
  const float2 tot = pair_sum(s, ss, part);
  const float md = tot.x / (float)D;
  const float mean = pivot + md;
  const float rstd = rsqrtf(fmaxf(tot.y / (float)D - md * md, 0.f) + eps);
  const float nmr = -mean * rstd;
  const int b = row / rows_per_batch;
  uint4* yr = reinterpret_cast<uint4*>(y + (size_t)row * ldy);
  const float* pa = scale ? scale + (size_t)b * mod_stride : weight;  // multiplier: (1 + scale) / scale / affine weight
  const float* pb = scale ? shift + (size_t)b * mod_stride : bias;
  constexpr int G = 2;  // vectors per batch: their parameter loads are in flight together
#pragma unroll
  for (int i0 = 0; i0 < HV; i0 += G) {
    float4 ta[G][2], tb[G][2];
    if (pa) {
#pragma unroll
      for (int k = 0; k < G; ++k) {
        const int idx = lane64 + (i0 + k) * 64;
        if (i0 + k < HV && idx < nvec) {
          const float4* qa = reinterpret_cast<const float4*>(pa + idx * 8);
          const float4* qb = reinterpret_cast<const float4*>(pb + idx * 8);
          ta[k][0] = qa[0]; ta[k][1] = qa[1];
          tb[k][0] = qb[0]; tb[k][1] = qb[1];
        }
      }
    }
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const int idx = lane64 + (i0 + k) * 64;
      if (i0 + k < HV && idx < nvec) {
        float f[8], o[8];
        unpack8(v[i0 + k], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(f[j], rstd, nmr);  // (x - mean) * rstd
        if (pa) {
          const float av[8] = {ta[k][0].x, ta[k][0].y, ta[k][0].z, ta[k][0].w, ta[k][1].x, ta[k][1].y, ta[k][1].z, ta[k][1].w};
          const float bv[8] = {tb[k][0].x, tb[k][0].y, tb[k][0].z, tb[k][0].w, tb[k][1].x, tb[k][1].y, tb[k][1].z, tb[k][1].w};
          if (scale && !scale_is_1p) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], 1.0f + av[j], bv[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], av[j], bv[j]);
          }
        }
        if (live) yr[idx] = pack8(o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int HV>
__global__ void __launch_bounds__(ROW_THREADS, 2)
rmsnorm_rope_kernel(bf16* __restrict__ x, int ldx, int rows, int D, float eps, const bf16* __restrict__ weight,
                    const float* __restrict__ rope_cos, const float* __restrict__ rope_sin, int L, int head_dim) {
  __shared__ float2 part[ROW_THREADS / 32];
  const int lane64 = threadIdx.x & 63;
  int row = blockIdx.x * (ROW_THREADS / 64) + (threadIdx.x >> 6);
  const bool live = row < rows;
  if (!live) row = rows - 1;
  const int nvec = D >> 3;
  uint4* xr = reinterpret_cast<uint4*>(x + (size_t)row * ldx);
  const uint4* wr = reinterpret_cast<const uint4*>(weight);
  uint4 v[HV];
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    const int idx = lane64 + i * 64;
    v[i] = idx < nvec ? xr[idx] : make_uint4(0u, 0u, 0u, 0u);
  }
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
  }
  const float2 tot = pair_sum(ss, 0.f, part);
  const float rstd = rsqrtf(tot.x / (float)D + eps);
  const int tok = rope_cos ? row % L : 0;
  const int half = head_dim >> 1;
#pragma unroll
  for (int i = 0; i < HV; ++i) {
    const int idx = lane64 + i * 64;
    if (idx < nvec) {
      float f[8], w[8], o[8];
      unpack8(v[i], f);
      unpack8(wr[idx], w);
      // diffusers RMSNorm with a bf16 weight: (x * rstd) rounded to bf16, then * weight (bf16 multiply)
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = bf16_round(bf16_round(f[j] * rstd) * w[j]);
      if (rope_cos) {
        const int c = idx * 8;
        const int pair0 = (c % head_dim) >> 1;  // 4 consecutive (even, odd) pairs of one head
        const float4 cs = *reinterpret_cast<const float4*>(rope_cos + (size_t)tok * half + pair0);
        const float4 sn = *reinterpret_cast<const float4*>(rope_sin + (size_t)tok * half + pair0);
        const float cv[4] = {cs.x, cs.y, cs.z, cs.w};
        const float sv[4] = {sn.x, sn.y, sn.z, sn.w};
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const float re = o[2 * p], im = o[2 * p + 1];
          o[2 * p] = re * cv[p] - im * sv[p];
          o[2 * p + 1] = re * sv[p] + im * cv[p];
        }
      }
      if (live) xr[idx] = pack8(o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
__global__ void patchify_kernel(const bf16* __restrict__ x, bf16* __restrict__ patches, int B, int C, int T, int H, int W) {
  const int hp = H >> 1, wp = W >> 1;
  const int K = C * 4;
  const size_t total = (size_t)B * T * hp * wp * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    size_t r = i / K;
    const int j = (int)(r % wp); r /= wp;
    const int ii = (int)(r % hp); r /= hp;
    const int f = (int)(r % T);
    const int b = (int)(r / T);
    const int c = k >> 2, dh = (k >> 1) & 1, dw = k & 1;
    patches[i] = x[((((size_t)b * C + c) * T + f) * H + (2 * ii + dh)) * W + (2 * j + dw)];
  }
}

__global__ void unpatchify_kernel(const bf16* __restrict__ y, int ldy, bf16* __restrict__ out, int B, int C, int T, int H, int W) {
  const int hp = H >> 1, wp = W >> 1;
  const size_t total = (size_t)B * C * T * H * W;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    size_t r = i;
    const int w = (int)(r % W); r /= W;
    const int hh = (int)(r % H); r /= H;
    const int f = (int)(r % T); r /= T;
    const int c = (int)(r % C);
    const int b = (int)(r / C);
    const size_t row = (((size_t)b * T + f) * hp + (hh >> 1)) * wp + (w >> 1);
    out[i] = y[row * ldy + ((hh & 1) * 2 + (w & 1)) * C + c];
  }
}

// ---------------------------------------------------------------------------------------------
constexpr int SL_MAX_B = 8;
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

template <bool W_BF16>
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ x, int K, const void* __restrict__ Wv, const void* __restrict__ biasv, int N,
                    int B, int act, int in_silu_bf16, float* __restrict__ out_f32, bf16* __restrict__ out_bf16) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= N) return;
  float acc[SL_MAX_B];
#pragma unroll
  for (int b = 0; b < SL_MAX_B; ++b) acc[b] = 0.f;
  for (int k = lane * 4; k < K; k += 128) {
    float w[4];
    if (W_BF16) {
      const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(Wv) + (size_t)warp * K + k);
      const float2 a = unpack_bf16x2(u.x), c = unpack_bf16x2(u.y);
      w[0] = a.x; w[1] = a.y; w[2] = c.x; w[3] = c.y;
    } else {
      const float4 u = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(Wv) + (size_t)warp * K + k);
      w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
    }
#pragma unroll
    for (int b = 0; b < SL_MAX_B; ++b) {
      if (b < B) {
        float4 xv = *reinterpret_cast<const float4*>(x + (size_t)b * K + k);
        if (in_silu_bf16) {  // act_fn(temb) evaluated on the bf16 tensor: SiLU in fp32, rounded to bf16
          xv.x = bf16_round(silu_f(xv.x)); xv.y = bf16_round(silu_f(xv.y));
          xv.z = bf16_round(silu_f(xv.z)); xv.w = bf16_round(silu_f(xv.w));
        }
        acc[b] += xv.x * w[0] + xv.y * w[1] + xv.z * w[2] + xv.w * w[3];
      }
    }
  }
#pragma unroll
  for (int b = 0; b < SL_MAX_B; ++b) acc[b] = warp_sum(acc[b]);
  if (lane == 0) {
    float bias = 0.f;
    if (biasv) bias = W_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(biasv)[warp]) : reinterpret_cast<const float*>(biasv)[warp];
    for (int b = 0; b < B; ++b) {
      float y = acc[b] + bias;
      if (out_bf16) y = bf16_round(y);  // the Linear output is a bf16 tensor in the reference
      if (act == 1) y = silu_f(y);
      if (out_f32) out_f32[(size_t)b * N + warp] = y;
      if (out_bf16) out_bf16[(size_t)b * N + warp] = __float2bfloat16_rn(y);
    }
  }
}

__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, float* __restrict__ emb, int B, int dim) {
  const int half = dim >> 1;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * half) return;
  const int b = i / half, k = i % half;
  // exponent = -ln(10000) * k / half (fp32, as torch.arange(float32) * scalar / half), freq = exp(exponent)
  const float exponent = (-9.210340371976184f * (float)k) / (float)half;
  const float arg = t[b] * expf(exponent);
  emb[(size_t)b * dim + k] = cosf(arg);
  emb[(size_t)b * dim + half + k] = sinf(arg);
}

__global__ void add_table_kernel(const float* __restrict__ table, int table_rows, const bf16* __restrict__ src, int src_ld,
                                 int src_per_chunk, float* __restrict__ dst, int B, int n, int chunks, unsigned plus_one_mask) {
  // dst[l, b, c, d] = table[l, c, d] + float(src[b, (src_per_chunk ? c*n : 0) + d])
  const size_t total = (size_t)table_rows * B * chunks * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int d = (int)(i % n);
    size_t r = i / n;
    const int c = (int)(r % chunks); r /= chunks;
    const int b = (int)(r % B);
    const int l = (int)(r / B);
    const float v = table[((size_t)l * chunks + c) * n + d] + __bfloat162float(src[(size_t)b * src_ld + (src_per_chunk ? c * n : 0) + d]);
    dst[i] = ((plus_one_mask >> c) & 1u) ? 1.0f + v : v;  // "scale" chunks are stored as (1 + scale)
  }
}

}  // namespace

int launch_layernorm(const bf16* x, int ldx, bf16* y, int ldy, int rows, int D, float eps, const float* scale,
                     const float* shift, int mod_stride, int rows_per_batch, const float* weight, const float* bias,
                     cudaStream_t stream, int scale_is_1p) {
  CE_REQUIRE(rows > 0 && D % 8 == 0 && D <= MAX_D, "layernorm: D must be a multiple of 8 and <= 8192");
  CE_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0, "layernorm: leading dims % 8");
  CE_REQUIRE((scale == nullptr) == (shift == nullptr), "layernorm: scale and shift come together");
  CE_REQUIRE((weight == nullptr) == (bias == nullptr), "layernorm: weight and bias come together");
  if (rows_per_batch <= 0) rows_per_batch = rows;
  const int grid = (rows + ROW_THREADS / 64 - 1) / (ROW_THREADS / 64);
  const int nvec = D / 8;
#define CE_LN(V) layernorm_kernel<V><<<grid, ROW_THREADS, 0, stream>>>(x, ldx, y, ldy, rows, D, eps, scale, shift, mod_stride, rows_per_batch, weight, bias, scale_is_1p)
  if (nvec <= 128) CE_LN(2);
  else if (nvec <= 256) CE_LN(4);
  else if (nvec <= 640) CE_LN(10);
  else CE_LN(16);
#undef CE_LN
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_rmsnorm_rope(bf16* x, int ldx, int rows, int D, float eps, const bf16* weight, const float* rope_cos,
                        const float* rope_sin, int L, int head_dim, cudaStream_t stream) {
  CE_REQUIRE(rows > 0 && D % 8 == 0 && D <= MAX_D, "rmsnorm: D must be a multiple of 8 and <= 8192");
  CE_REQUIRE(ldx % 8 == 0 && weight != nullptr, "rmsnorm: ldx % 8, weight");
  if (rope_cos) CE_REQUIRE(rope_sin && L > 0 && head_dim % 8 == 0 && D % head_dim == 0, "rmsnorm: rope table / head_dim");
  const int grid = (rows + ROW_THREADS / 64 - 1) / (ROW_THREADS / 64);
  const int nvec = D / 8;
#define CE_RMS(V) rmsnorm_rope_kernel<V><<<grid, ROW_THREADS, 0, stream>>>(x, ldx, rows, D, eps, weight, rope_cos, rope_sin, L, head_dim)
  if (nvec <= 128) CE_RMS(2);
  else if (nvec <= 256) CE_RMS(4);
  else if (nvec <= 640) CE_RMS(10);
  else CE_RMS(16);
#undef CE_RMS
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_layernorm_stats(const bf16* x, int ldx, bf16* y, int ldy, int rows, int D, float eps, const float* scale, const float* shift,
                           int mod_stride, int rows_per_batch, const float* weight, const float* bias, int scale_is_1p, const float2* stats,
                           int stats_ld, int tiles, int tile_n, cudaStream_t stream) {
  CE_REQUIRE(rows > 0 && D % 8 == 0 && D <= MAX_D && ldx % 8 == 0 && ldy % 8 == 0, "layernorm(stats): shapes");
  CE_REQUIRE(stats != nullptr && tiles >= 1 && tiles <= 32 && stats_ld >= tiles && tile_n > 0 && (tiles - 1) * tile_n < D && tiles * tile_n >= D,
             "layernorm(stats): the tile partials must cover exactly [0, D) in at most 32 tiles");
  CE_REQUIRE((scale == nullptr) == (shift == nullptr) && (weight == nullptr) == (bias == nullptr), "layernorm(stats): parameter pairs");
  if (rows_per_batch <= 0) rows_per_batch = rows;
  const int grid = (rows + ROW_THREADS / 64 - 1) / (ROW_THREADS / 64);
  const int nvec = D / 8;
#define CE_LNS(V) layernorm_stats_kernel<V><<<grid, ROW_THREADS, 0, stream>>>(x, ldx, y, ldy, rows, D, eps, scale, shift, mod_stride, rows_per_batch, weight, bias, scale_is_1p, stats, stats_ld, tiles, tile_n)
  if (nvec <= 128) CE_LNS(2);
  else if (nvec <= 256) CE_LNS(4);
  else if (nvec <= 640) CE_LNS(10);
  else CE_LNS(16);
#undef CE_LNS
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_rmsnorm_rope_stats(bf16* x, int ldx, int rows, int D, float eps, const bf16* weight0, const bf16* weight1, int nmat,
                              const float* rope_cos, const float* rope_sin, int L, int head_dim, const float2* stats, int stats_ld, int tiles,
                              cudaStream_t stream, const SpScatter* sp) {
  CE_REQUIRE(rows > 0 && D % 8 == 0 && D <= MAX_D && ldx % 8 == 0 && weight0 != nullptr, "rmsnorm(stats): shapes");
  CE_REQUIRE(nmat == 1 || (nmat == 2 && weight1 != nullptr), "rmsnorm(stats): one or two matrices");
  CE_REQUIRE(stats != nullptr && tiles >= 1 && tiles <= 32 && stats_ld >= nmat * tiles, "rmsnorm(stats): tile partials");
  if (rope_cos) CE_REQUIRE(rope_sin && L > 0 && head_dim % 8 == 0 && D % head_dim == 0, "rmsnorm(stats): rope table / head_dim");
  dim3 grid((rows + ROW_THREADS / 64 - 1) / (ROW_THREADS / 64), nmat);
  const int nvec = D / 8;
  SpScatter spv;
  if (sp) spv = *sp;
  if (spv.world) CE_REQUIRE(spv.world <= 8 && spv.heads_per_rank > 0 && spv.rows_per_batch > 0 && rows % spv.rows_per_batch == 0, "rmsnorm(stats): scatter");
#define CE_RMSS(V) rmsnorm_rope_stats_kernel<V><<<grid, ROW_THREADS, 0, stream>>>(x, ldx, rows, D, eps, weight0, weight1, rope_cos, rope_sin, L, head_dim, stats, stats_ld, tiles, spv)
  if (nvec <= 128) CE_RMSS(2);
  else if (nvec <= 256) CE_RMSS(4);
  else if (nvec <= 640) CE_RMSS(10);
  else CE_RMSS(16);
#undef CE_RMSS
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_patchify(const bf16* x, bf16* patches, int B, int C, int T, int H, int W, cudaStream_t stream) {
  CE_REQUIRE(H % 2 == 0 && W % 2 == 0, "patchify: H, W must be even");
  const size_t total = (size_t)B * T * (H / 2) * (W / 2) * C * 4;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  patchify_kernel<<<grid, 256, 0, stream>>>(x, patches, B, C, T, H, W);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_unpatchify(const bf16* y, int ldy, bf16* out, int B, int C, int T, int H, int W, cudaStream_t stream) {
  const size_t total = (size_t)B * C * T * H * W;
  const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  unpatchify_kernel<<<grid, 256, 0, stream>>>(y, ldy, out, B, C, T, H, W);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_small_linear(const float* x, int K, const void* W, const void* bias, int w_is_bf16, int N, int B, int act,
                        int in_silu_bf16, float* out_f32, bf16* out_bf16, cudaStream_t stream) {
  CE_REQUIRE(B >= 1 && B <= SL_MAX_B, "small_linear: batch must be 1..8");
  CE_REQUIRE(K % 4 == 0, "small_linear: K % 4");
  const int blocks = (N * 32 + 255) / 256;
  if (w_is_bf16)
    small_linear_kernel<true><<<blocks, 256, 0, stream>>>(x, K, W, bias, N, B, act, in_silu_bf16, out_f32, out_bf16);
  else
    small_linear_kernel<false><<<blocks, 256, 0, stream>>>(x, K, W, bias, N, B, act, in_silu_bf16, out_f32, out_bf16);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_timestep_sinusoid(const float* t, float* emb, int B, int dim, cudaStream_t stream) {
  const int n = B * (dim / 2);
  timestep_sinusoid_kernel<<<(n + 127) / 128, 128, 0, stream>>>(t, emb, B, dim);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_add_table(const float* table, int table_rows, const bf16* src, int src_ld, float* dst, int B, int n,
                     int chunks, cudaStream_t stream, unsigned plus_one_mask) {
  const size_t total = (size_t)table_rows * B * chunks * n;
  const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  const int src_per_chunk = src_ld >= chunks * n ? 1 : 0;
  add_table_kernel<<<grid, 256, 0, stream>>>(table, table_rows, src, src_ld, src_per_chunk, dst, B, n, chunks, plus_one_mask);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
