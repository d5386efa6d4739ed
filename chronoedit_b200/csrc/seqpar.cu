#include "seqpar.cuh"

#include <string.h>

#include "../../include/chronoedit_b200.h"

namespace ce {

namespace {

struct Ptrs8 {
  void* p[8];
};

__global__ void sp_barrier_kernel(Ptrs8 flags, int rank, int world, uint32_t epoch) {
  const int w = threadIdx.x;
  if (w >= world) return;
  __threadfence_system();
  uint32_t* remote = reinterpret_cast<uint32_t*>(flags.p[w]) + rank;   // my slot in rank w's flag array
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[rank]) + w;   // rank w's slot in my array
  const uint64_t t0 = global_timer_ns();
  uint32_t v = 0, spins = 0;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if ((++spins & 0x3FFF) == 0 && global_timer_ns() - t0 > 20000000000ull) {
      printf("[chronoedit_b200] sequence-parallel barrier timeout: rank %d waits for rank %d (epoch %u, saw %u)\n", rank, w, epoch, v);
      __trap();
    }
  } while ((int32_t)(v - epoch) < 0);
  __threadfence_system();
}

__global__ void sp_scatter_cols_kernel(const bf16* __restrict__ src, int ld_src, int B, int rows_per_batch, int ncols, Ptrs8 dst, int cols_per_rank,
                                       int dst_rows_per_batch, int dst_row0) {
  const int nv = ncols / 8;
  const size_t total = (size_t)B * rows_per_batch * nv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % nv) * 8;
    const size_t r = i / nv;
    const int li = (int)(r % rows_per_batch);
    const int b = (int)(r / rows_per_batch);
    const uint4 v = *reinterpret_cast<const uint4*>(src + r * ld_src + c);
    bf16* d = reinterpret_cast<bf16*>(dst.p[c / cols_per_rank]) + ((size_t)b * dst_rows_per_batch + dst_row0 + li) * cols_per_rank + c % cols_per_rank;
    *reinterpret_cast<uint4*>(d) = v;
  }
}

__global__ void sp_broadcast_rows_kernel(const bf16* __restrict__ src, int B, int rows_per_batch, int ncols, Ptrs8 dst, int world, int dst_rows_per_batch,
                                         int dst_row0) {
  const int nv = ncols / 8;
  const size_t total = (size_t)B * rows_per_batch * nv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % nv) * 8;
    const size_t r = i / nv;
    const int li = (int)(r % rows_per_batch);
    const int b = (int)(r / rows_per_batch);
    const uint4 v = *reinterpret_cast<const uint4*>(src + r * ncols + c);
    for (int w = 0; w < world; ++w)
      *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(dst.p[w]) + ((size_t)b * dst_rows_per_batch + dst_row0 + li) * ncols + c) = v;
  }
}

__global__ void patchify_range_kernel(const bf16* __restrict__ x, bf16* __restrict__ patches, int B, int C, int T, int H, int W, int tok0, int ntok) {
  const int hp = H >> 1, wp = W >> 1;
  const int K = C * 4;
  const size_t total = (size_t)B * ntok * K;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    size_t r = i / K;
    const int tok = tok0 + (int)(r % ntok);
    const int b = (int)(r / ntok);
    const int j = tok % wp, ii = (tok / wp) % hp, f = tok / (wp * hp);
    const int c = k >> 2, dh = (k >> 1) & 1, dw = k & 1;
    patches[i] = x[((((size_t)b * C + c) * T + f) * H + (2 * ii + dh)) * W + (2 * j + dw)];
  }
}

inline int grid_for(size_t n) {
  size_t g = (n + 255) / 256;
  return (int)(g < 148 * 16 ? (g ? g : 1) : 148 * 16);
}

}  // namespace

SpLayout sp_layout(int64_t B, int64_t L, int64_t D, int64_t No, int world) {
  auto al = [](int64_t v) { return (v + 1023) & ~int64_t(1023); };
  SpLayout l;
  const int64_t g = al(B * L * (D / world) * 2);
  l.q = 0;
  l.k = g;
  l.v = 2 * g;
  l.attn = 3 * g;
  l.yout = l.attn + al(B * (L / world) * D * 2);
  l.flags = l.yout + al(B * L * No * 2);
  l.bytes = l.flags + 1024;
  return l;
}

int launch_sp_barrier(SeqPar& sp, const SpLayout& lay, cudaStream_t stream) {
  Ptrs8 f;
  for (int w = 0; w < 8; ++w) f.p[w] = w < sp.world ? sp.region[w] + lay.flags : nullptr;
  ++sp.epoch;
  sp_barrier_kernel<<<1, 32, 0, stream>>>(f, sp.rank, sp.world, sp.epoch);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_sp_scatter_cols(const bf16* src, int ld_src, int B, int rows_per_batch, int ncols, bf16* const* dst, int world, int cols_per_rank,
                           int dst_rows_per_batch, int dst_row0, cudaStream_t stream) {
  CE_REQUIRE(ncols % 8 == 0 && cols_per_rank % 8 == 0 && ld_src % 8 == 0 && world <= 8 && cols_per_rank * world == ncols, "sp scatter: shapes");
  Ptrs8 d;
  for (int w = 0; w < 8; ++w) d.p[w] = w < world ? dst[w] : nullptr;
  sp_scatter_cols_kernel<<<grid_for((size_t)B * rows_per_batch * ncols / 8), 256, 0, stream>>>(src, ld_src, B, rows_per_batch, ncols, d, cols_per_rank,
                                                                                               dst_rows_per_batch, dst_row0);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_sp_broadcast_rows(const bf16* src, int B, int rows_per_batch, int ncols, bf16* const* dst, int world, int dst_rows_per_batch,
                             int dst_row0, cudaStream_t stream) {
  CE_REQUIRE(ncols % 8 == 0 && world <= 8, "sp broadcast: shapes");
  Ptrs8 d;
  for (int w = 0; w < 8; ++w) d.p[w] = w < world ? dst[w] : nullptr;
  sp_broadcast_rows_kernel<<<grid_for((size_t)B * rows_per_batch * ncols / 8), 256, 0, stream>>>(src, B, rows_per_batch, ncols, d, world,
                                                                                                 dst_rows_per_batch, dst_row0);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

int launch_patchify_range(const bf16* x, bf16* patches, int B, int C, int T, int H, int W, int tok0, int ntok, cudaStream_t stream) {
  CE_REQUIRE(H % 2 == 0 && W % 2 == 0 && ntok > 0, "patchify(range): shapes");
  patchify_range_kernel<<<grid_for((size_t)B * ntok * C * 4), 256, 0, stream>>>(x, patches, B, C, T, H, W, tok0, ntok);
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}

}  // namespace ce

using namespace ce;

extern "C" {

// ---- CUDA IPC plumbing for the peer regions (cudaMalloc'ed here so that the IPC handle covers exactly the region)
int ce_ipc_alloc(int64_t bytes, void** ptr) {
  CE_REQUIRE(bytes > 0 && ptr, "ce_ipc_alloc: arguments");
  CE_CHECK_CUDA(cudaMalloc(ptr, (size_t)bytes));
  CE_CHECK_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  CE_CHECK_CUDA(cudaDeviceSynchronize());
  return CE_OK;
}
int ce_ipc_free(void* ptr) {
  if (ptr) CE_CHECK_CUDA(cudaFree(ptr));
  return CE_OK;
}
int ce_ipc_get_handle(void* ptr, void* handle_out_64_bytes) {
  CE_REQUIRE(ptr && handle_out_64_bytes, "ce_ipc_get_handle: arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CE_CHECK_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle_out_64_bytes), ptr));
  return CE_OK;
}
int ce_ipc_open(const void* handle_64_bytes, void** ptr) {
  CE_REQUIRE(handle_64_bytes && ptr, "ce_ipc_open: arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64_bytes, 64);
  CE_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return CE_OK;
}
int ce_ipc_close(void* ptr) {
  if (ptr) CE_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return CE_OK;
}

}  // extern "C"
