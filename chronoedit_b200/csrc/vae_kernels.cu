// HBM-bound kernels of the Wan VAE path (channels-last bf16): RMS_norm(+SiLU), nearest 2x upsample, im2row for the
// few small-Cin convolutions, row softmax for the single-head mid attention, layout conversion.
#include "vae_kernels.cuh"

namespace ce {

namespace {

__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }

// RMS_norm.forward (wan2pt1.py:74-75) evaluated with the reference's bf16 rounding points:
//   n = bf16(||x||_2) ; y = bf16(x / max(n, 1e-12)) ; y = bf16(y * sqrt(C)) ; y = bf16(y * gamma) ; [y = bf16(silu(y))]
// One pixel per group of G lanes (G = 16 for C <= 128, else 32); 16-byte accesses.
template <int G>
__global__ void __launch_bounds__(256)
rms_silu_cl_kernel(const bf16* __restrict__ x, const bf16* __restrict__ gamma, bf16* __restrict__ y, size_t pixels, int C,
                   int apply_silu) {
  const int lane_in_group = threadIdx.x % G;
  const size_t group = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const size_t ngroups = ((size_t)gridDim.x * blockDim.x) / G;
  const int nvec = C >> 3;
  const float sqrt_c = sqrtf((float)C);
  for (size_t p = group; p < pixels; p += ngroups) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + p * C);
    uint4 v[2];
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = lane_in_group + i * G;
      if (idx < nvec) {
        v[i] = xr[idx];
        const float2 a = unpack_bf16x2(v[i].x), b = unpack_bf16x2(v[i].y), c = unpack_bf16x2(v[i].z), d = unpack_bf16x2(v[i].w);
        ss += a.x * a.x + a.y * a.y + b.x * b.x + b.y * b.y + c.x * c.x + c.y * c.y + d.x * d.x + d.y * d.y;
      }
    }
    const uint32_t gmask = (G == 32) ? 0xffffffffu : (0xFFFFu << ((threadIdx.x & 31) & 16));
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(gmask, ss, o, G);
    const float n = fmaxf(bf16_round(sqrtf(ss)), 1e-12f);
    uint4* yr = reinterpret_cast<uint4*>(y + p * C);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int idx = lane_in_group + i * G;
      if (idx < nvec) {
        const uint4 gv = reinterpret_cast<const uint4*>(gamma)[idx];
        const uint32_t xin[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
        const uint32_t gin[4] = {gv.x, gv.y, gv.z, gv.w};
        uint32_t out[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 xv = unpack_bf16x2(xin[k]);
          const float2 gg = unpack_bf16x2(gin[k]);
          float a = bf16_round(bf16_round(bf16_round(xv.x / n) * sqrt_c) * gg.x);
          float b = bf16_round(bf16_round(bf16_round(xv.y / n) * sqrt_c) * gg.y);
          if (apply_silu) {
            a = silu_f(a);
            b = silu_f(b);
          }
          out[k] = pack_bf16x2(a, b);
        }
        yr[idx] = make_uint4(out[0], out[1], out[2], out[3]);
      }
    }
  }
}

// out[t, 2h+a, 2w+b, :] = in[t, h, w, :]   (nn.Upsample(scale 2, nearest-exact), wan2pt1.py:78-83, 98-104)
__global__ void upsample2x_cl_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int T, int H, int W, int cvec) {
  const size_t total = (size_t)T * (2 * H) * (2 * W) * cvec;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cvec);
    size_t r = i / cvec;
    const int ow = (int)(r % (2 * W)); r /= (2 * W);
    const int oh = (int)(r % (2 * H));
    const int t = (int)(r / (2 * H));
    out[i] = in[(((size_t)t * H + (oh >> 1)) * W + (ow >> 1)) * cvec + c];
  }
}

// A[(to,oh,ow), tap*Cin + c] = X[t_base + to + dt, oh + dh - ph, ow + dw - pw, c]  (0 outside / in the K padding).
// X is addressed through explicit strides so planar (NCTHW) and channels-last inputs both work.
// One thread = one output pixel x eight consecutive k: a single 16-byte store (the matrix is 236 MB per 720p frame and this kernel is
// pure HBM write traffic), 32-bit index arithmetic, the (at most 81) scattered input reads hit L1/L2.
__global__ void __launch_bounds__(256)
im2row_kernel(const bf16* __restrict__ x, size_t sc, size_t st, size_t sh, size_t sw, int Tin, int Hin, int Win, int Cin,
              bf16* __restrict__ A, int Kpad, int Tout, int Hout, int Wout, int kt, int kh, int kw, int ph, int pw, int t_base) {
  const int groups = Kpad >> 3;
  const size_t total = (size_t)Tout * Hout * Wout * groups;
  const int K = kt * kh * kw * Cin;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int g = (int)(i % groups);
    uint32_t r = (uint32_t)(i / groups);   // pixel index < 2^31 (checked by the launcher)
    const int ow = (int)(r % (uint32_t)Wout);
    r /= (uint32_t)Wout;
    const int oh = (int)(r % (uint32_t)Hout);
    const int to = (int)(r / (uint32_t)Hout);
    uint32_t w[4];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = g * 8 + e;
      unsigned short bits = 0;
      if (k < K) {
        const int c = k % Cin;
        int tap = k / Cin;
        const int dw = tap % kw;
        tap /= kw;
        const int dh = tap % kh;
        const int dt = tap / kh;
        const int ti = t_base + to + dt, hi = oh + dh - ph, wi = ow + dw - pw;
        if (ti >= 0 && ti < Tin && hi >= 0 && hi < Hin && wi >= 0 && wi < Win)
          bits = __bfloat16_as_ushort(x[c * sc + ti * st + hi * sh + wi * sw]);
      }
      if (e & 1) w[e >> 1] |= (uint32_t)bits << 16;
      else w[e >> 1] = bits;
    }
    *reinterpret_cast<uint4*>(A + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// P[r, :] = softmax(S[r, :] * scale) in fp32, stored bf16.
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ S, int lds, bf16* __restrict__ P, int ldp, int cols, float scale) {
  __shared__ float red[32];
  const float* s = S + (size_t)blockIdx.x * lds;
  bf16* p = P + (size_t)blockIdx.x * ldp;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) mx = fmaxf(mx, s[i]);
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  mx = warp_max((threadIdx.x & 31) < (blockDim.x >> 5) ? red[threadIdx.x & 31] : -INFINITY);
  __syncthreads();
  float sum = 0.f;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) sum += __expf((s[i] - mx) * scale);
  sum = warp_sum(sum);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
  __syncthreads();
  sum = warp_sum((threadIdx.x & 31) < (blockDim.x >> 5) ? red[threadIdx.x & 31] : 0.f);
  const float inv = 1.0f / sum;
  for (int i = threadIdx.x; i < cols; i += blockDim.x) p[i] = __float2bfloat16_rn(__expf((s[i] - mx) * scale) * inv);
}

// planar[c, p] = cl[p, c] for c < Cout (posterior mean = first z_dim channels of conv1's output)
__global__ void cl_to_planar_kernel(const bf16* __restrict__ cl, int C, bf16* __restrict__ planar, size_t pixels, int Cout) {
  const size_t total = pixels * Cout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t p = i % pixels;
    const int c = (int)(i / pixels);
    planar[i] = cl[p * C + c];
  }
}

int grid_for(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = (size_t)148 * 16;
  return (int)(g < cap ? (g == 0 ? 1 : g) : cap);
}

}  // namespace

int launch_rms_silu_cl(const bf16* x, const bf16* gamma, bf16* y, size_t pixels, int C, int apply_silu, cudaStream_t stream) {
  CE_REQUIRE(C % 8 == 0 && C <= 512, "rms_silu: C % 8, C <= 512");
  if (C <= 256 / 2) {
    const size_t threads = pixels * 16;
    rms_silu_cl_kernel<16><<<grid_for(threads, 256), 256, 0, stream>>>(x, gamma, y, pixels, C, apply_silu);
  } else {
    const size_t threads = pixels * 32;
    rms_silu_cl_kernel<32><<<grid_for(threads, 256), 256, 0, stream>>>(x, gamma, y, pixels, C, apply_silu);
  }
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_upsample2x_cl(const bf16* in, bf16* out, int T, int H, int W, int C, cudaStream_t stream) {
  CE_REQUIRE(C % 8 == 0, "upsample: C % 8");
  const size_t total = (size_t)T * 4 * H * W * (C / 8);
  upsample2x_cl_kernel<<<grid_for(total, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(in), reinterpret_cast<uint4*>(out), T, H, W, C / 8);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_im2row(const bf16* x, size_t sc, size_t st, size_t sh, size_t sw, int Tin, int Hin, int Win, int Cin, bf16* A, int Kpad,
                  int Tout, int Hout, int Wout, int kt, int kh, int kw, int ph, int pw, int t_base, cudaStream_t stream) {
  CE_REQUIRE(Kpad % 8 == 0 && Kpad >= kt * kh * kw * Cin, "im2row: Kpad");
  CE_REQUIRE((size_t)Tout * Hout * Wout < (size_t(1) << 31) && (reinterpret_cast<uintptr_t>(A) & 15) == 0, "im2row: pixel count / alignment");
  const size_t total = (size_t)Tout * Hout * Wout * (Kpad / 8);
  im2row_kernel<<<grid_for(total, 256), 256, 0, stream>>>(x, sc, st, sh, sw, Tin, Hin, Win, Cin, A, Kpad, Tout, Hout, Wout, kt, kh, kw, ph, pw, t_base);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_softmax_rows(const float* S, int lds, bf16* P, int ldp, int rows, int cols, float scale, cudaStream_t stream) {
  softmax_rows_kernel<<<rows, 256, 0, stream>>>(S, lds, P, ldp, cols, scale);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_cl_to_planar(const bf16* cl, int C, bf16* planar, size_t pixels, int Cout, cudaStream_t stream) {
  cl_to_planar_kernel<<<grid_for(pixels * Cout, 256), 256, 0, stream>>>(cl, C, planar, pixels, Cout);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce
