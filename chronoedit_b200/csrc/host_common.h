// Host-side helpers shared by all translation units: status/error plumbing of the C ABI and
// CUtensorMap construction (driver entry point fetched at run time, so the library has no
// link-time dependency on libcuda and builds on a GPU-less box).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace ce {

// ---- error plumbing ------------------------------------------------------------------------
// Every extern "C" entry point returns 0 on success or a negative code and records a message
// retrievable with ce_last_error().  No exceptions cross the ABI.
enum Status : int {
  CE_OK = 0,
  CE_ERR_INVALID = -1,      // bad argument / unsupported shape
  CE_ERR_CUDA = -2,         // CUDA runtime / driver error
  CE_ERR_NO_DEVICE = -3,    // no sm_100 device: there is NO CPU fallback
  CE_ERR_MISSING_WEIGHT = -4,
  CE_ERR_WORKSPACE = -5,
};

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);

#define CE_CHECK_CUDA(expr)                                                                       \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess)                                                                        \
      return ::ce::fail(::ce::CE_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));   \
  } while (0)

#define CE_REQUIRE(cond, msg)                                                        \
  do {                                                                               \
    if (!(cond)) return ::ce::fail(::ce::CE_ERR_INVALID, std::string(msg) + " [" #cond "]"); \
  } while (0)

// ---- tensor maps ---------------------------------------------------------------------------
// bf16 tensor, `rank` dims listed innermost first.  strides_bytes[i] is the byte stride of dim i+1
// (dim 0 is contiguous).  128-byte swizzle, zero OOB fill.
// elem_strides (optional): traversal stride per dim; TMA then fetches ceil(box[i] / elem_strides[i]) elements along dim i.
int make_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                   const uint32_t* box, const uint32_t* elem_strides = nullptr);

// 2-D row-major [rows, cols] bf16 with leading dimension `ld` (elements); box = [box_rows, 64 cols].
int make_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows);

int device_sm_count();   // of the CURRENT device (cached per device index)
int check_device();
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, and a process
// may drive several GPUs (VAE on cuda:1, DiT on cuda:0; model.to() after a first forward).
int ensure_dynamic_smem(const void* kernel, int bytes);
#define CE_ENSURE_SMEM(kernel, bytes)                                                              \
  do {                                                                                             \
    int _rc = ::ce::ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), (int)(bytes));      \
    if (_rc) return _rc;                                                                           \
  } while (0)  // CE_OK iff the current device is compute capability 10.x

}  // namespace ce
