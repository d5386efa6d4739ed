// Micro-benchmark: back-to-back tcgen05.mma issue rate per SM for the shapes the attention kernels use.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o mma_rate mma_rate.cu -lcuda && ./mma_rate
// mode 0: SS  M128 N64  K16 (Q.K^T with 64-key tiles)      mode 1: SS M128 N128 K16      mode 2: SS M128 N256 K16
// mode 3: TS  M128 N128 K16 (P in TMEM, V MN-major)        mode 4: alternate mode-0 and mode-3 groups (64-key-tile mix)
// mode 5: alternate mode-1 and mode-3 groups (attention2 mix)
#include <cstdio>
#include <cuda_runtime.h>
#include "../../chronoedit_b200/csrc/ptx.cuh"
using namespace ce;

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(long long* out, int rounds, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t slot;
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + (i & 7);
  if (warp == 0) tmem_alloc(&slot, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem);               // 128 x 128 bf16, two 64-wide halves of 16 KB
    const uint32_t b_addr = smem_u32(smem) + 32 * 1024;   // up to 256 x 128 bf16
    constexpr uint32_t I64 = umma_idesc_bf16(128, 64, 0), I128 = umma_idesc_bf16(128, 128, 0), I256 = umma_idesc_bf16(128, 256, 0);
    constexpr uint32_t IPV = umma_idesc_bf16(128, 128, 1);
    const long long t0 = clock64();
    for (int r = 0; r < rounds; ++r) {
      const int m = mode < 4 ? mode : ((r & 1) ? 3 : (mode == 4 ? 0 : 1));
      if (m == 0) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(base, umma_desc_kmajor_sw128(a_addr + (kk >> 2) * 16384) + 2 * (kk & 3), umma_desc_kmajor_sw128(b_addr + (kk >> 2) * 8192) + 2 * (kk & 3), I64, kk != 0);
      } else if (m == 1) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(base, umma_desc_kmajor_sw128(a_addr + (kk >> 2) * 16384) + 2 * (kk & 3), umma_desc_kmajor_sw128(b_addr + (kk >> 2) * 16384) + 2 * (kk & 3), I128, kk != 0);
      } else if (m == 2) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ss(base, umma_desc_kmajor_sw128(a_addr + (kk >> 2) * 16384) + 2 * (kk & 3), umma_desc_kmajor_sw128(b_addr + (kk >> 2) * 32768) + 2 * (kk & 3), I256, kk != 0);
      } else {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_bf16_ts(base + 256, base + 128 + kk * 8, umma_desc_mnmajor_sw128(b_addr + kk * 2048, 16384), IPV, kk != 0);
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0, 1);
    const long long t1 = clock64();
    out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(base, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  long long h[148];
  const int rounds = 2000;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 97 * 1024);
  const char* names[6] = {"SS M128 N64", "SS M128 N128", "SS M128 N256", "TS M128 N128 (V MN-major)", "mix SS N64 / TS N128", "mix SS N128 / TS N128"};
  const double ideal[6] = {32, 64, 128, 64, 48, 64};
  for (int mode = 0; mode < 6; ++mode) {
    for (int rep = 0; rep < 2; ++rep) {
      mma_rate_kernel<<<148, 128, 97 * 1024>>>(d, rounds, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        return 1;
      }
    }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double s = 0;
    for (int i = 0; i < 148; ++i) s += h[i];
    printf("%-28s %7.1f cycles per K=16 instruction (floor %.0f)\n", names[mode], s / 148 / (rounds * 8.0), ideal[mode]);
  }
  return 0;
}
