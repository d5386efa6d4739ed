// Per-step sampling glue around the DiT call, fused into ONE elementwise launch:
//   classifier-free-guidance combine  (chronoedit_diffusers/pipeline_chronoedit.py:736)
//   flow-prediction x0 conversion     (chronoedit/_src/models/fm_solvers_unipc.py:329-346)
//   UniC corrector, bh2, order 1 / 2  (:565-640)
//   UniP predictor, bh2, order 1 / 2  (:440-499)
//   next model input channels [0, c_lat) of cat([latents, condition], 1).to(bf16)   (pipeline_chronoedit.py:712)
// The reference runs ~25 separate torch kernels per step for this; every one of them rounds its result to the tensor dtype
// (bf16 in the diffusers pipeline, fp32 in the native loop).  The kernel keeps the same op order and rounds at the same
// points, so the result is bit-identical: r<T>() below is "what a torch elementwise kernel would have stored".  fp32 math
// uses the _rn intrinsics so that ptxas cannot contract a multiply and an add into an FMA (one rounding instead of two).
//
// HBM-bound and tiny (a 720p / 5-frame latent is 460 800 elements, <= 9 streams of it, ~8 MB): what matters is that it is
// one launch with no host synchronisation, not a roofline fraction.
#include <cuda_bf16.h>

#include "../../include/chronoedit_b200.h"
#include "host_common.h"

namespace ce {

namespace {

template <typename T>
struct Io;
template <>
struct Io<float> {
  __device__ static float round(float x) { return x; }
};
template <>
struct Io<__nv_bfloat16> {
  __device__ static float round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

template <typename T>
__device__ __forceinline__ float ld(const void* p, int64_t i) {
  if constexpr (sizeof(T) == 4) return reinterpret_cast<const float*>(p)[i];
  else return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
}
template <typename T>
__device__ __forceinline__ void st(void* p, int64_t i, float v) {
  if constexpr (sizeof(T) == 4) reinterpret_cast<float*>(p)[i] = v;
  else reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16_rn(v);
}

// S = sample / state dtype, V = model-output dtype.  Supported: (f32,f32) (f32,bf16) (bf16,bf16): the promoted dtype is S.
template <typename S, typename V>
__global__ void __launch_bounds__(256) unipc_step_kernel(const ce_unipc_step_args a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += stride) {
    auto rs = [](float x) { return Io<S>::round(x); };
    auto rv = [](float x) { return Io<V>::round(x); };
    // ---- model output (with guidance)
    float v = ld<V>(a.cond, i);
    if (a.uncond != nullptr) {
      const float u = ld<V>(a.uncond, i);
      v = rv(__fadd_rn(u, rv(__fmul_rn(a.guidance, rv(__fsub_rn(v, u))))));
    }
    // ---- x0 prediction
    float x = ld<S>(a.sample, i);
    const float m_t = rs(__fsub_rn(x, rv(__fmul_rn(a.sigma, v))));
    const float m_p = a.m_prev ? ld<S>(a.m_prev, i) : 0.f;
    // ---- corrector: re-derive the current sample from the previous one now that the model has seen it
    if (a.use_corrector) {
      const float xt = rs(__fsub_rn(rs(__fmul_rn(a.c_x, ld<S>(a.last_sample, i))), rs(__fmul_rn(a.c_m0, m_p))));
      const float dt = rs(__fsub_rn(m_t, m_p));
      float inner;
      if (a.c_order == 1) {
        inner = rs(__fadd_rn(0.f, rs(__fmul_rn(0.5f, dt))));
      } else {
        const float d1 = rs(__fmul_rn(rs(__fsub_rn(ld<S>(a.m_prev2, i), m_p)), a.c_inv_rk));
        inner = rs(__fadd_rn(rs(__fmul_rn(a.c_rho0, d1)), rs(__fmul_rn(a.c_rho1, dt))));
      }
      x = rs(__fsub_rn(xt, rs(__fmul_rn(a.c_bh, inner))));
      st<S>(a.corrected_out, i, x);
    }
    // ---- predictor
    const float xt = rs(__fsub_rn(rs(__fmul_rn(a.p_x, x)), rs(__fmul_rn(a.p_m0, m_t))));
    float nxt;
    if (a.p_order == 1) {
      nxt = rs(__fsub_rn(xt, a.p_zero));
    } else {
      const float d1 = rs(__fmul_rn(rs(__fsub_rn(m_p, m_t)), a.p_inv_rk));
      nxt = rs(__fsub_rn(xt, rs(__fmul_rn(a.p_bh, rs(__fmul_rn(0.5f, d1))))));
    }
    st<S>(a.x0_out, i, m_t);
    st<S>(a.prev_sample_out, i, nxt);
    if (a.model_input_out != nullptr) {
      const int64_t per_b = (int64_t)a.c_lat * a.inner;
      const int64_t b = i / per_b;
      reinterpret_cast<__nv_bfloat16*>(a.model_input_out)[b * (int64_t)a.c_total * a.inner + (i - b * per_b)] = __float2bfloat16_rn(nxt);
    }
  }
}

}  // namespace

}  // namespace ce

extern "C" int ce_unipc_step(const ce_unipc_step_args* a, void* stream) {
  using namespace ce;
  CE_REQUIRE(a != nullptr, "unipc_step: null args");
  CE_REQUIRE(a->n > 0, "unipc_step: empty latent");
  CE_REQUIRE(a->cond && a->sample && a->x0_out && a->prev_sample_out, "unipc_step: null tensor");
  CE_REQUIRE(a->p_order == 1 || a->p_order == 2, "unipc_step: predictor order must be 1 or 2 (solver_order 2)");
  CE_REQUIRE(a->p_order == 1 || a->m_prev != nullptr, "unipc_step: order-2 predictor needs the previous x0 prediction");
  if (a->use_corrector) {
    CE_REQUIRE(a->c_order == 1 || a->c_order == 2, "unipc_step: corrector order must be 1 or 2");
    CE_REQUIRE(a->last_sample && a->m_prev && a->corrected_out, "unipc_step: corrector needs last_sample, m_prev, corrected_out");
    CE_REQUIRE(a->c_order == 1 || a->m_prev2 != nullptr, "unipc_step: order-2 corrector needs two previous x0 predictions");
  }
  if (a->model_input_out) {
    CE_REQUIRE(a->inner > 0 && a->c_lat > 0 && a->c_total >= a->c_lat && a->n % ((int64_t)a->c_lat * a->inner) == 0,
               "unipc_step: model_input_out needs n = B * c_lat * inner");
  }
  int rc = check_device();
  if (rc != CE_OK) return rc;
  const int threads = 256;
  const int64_t want = (a->n + threads - 1) / threads;
  const int blocks = (int)(want < (int64_t)device_sm_count() * 8 ? want : (int64_t)device_sm_count() * 8);
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  if (a->sample_dtype == CE_DTYPE_F32 && a->model_dtype == CE_DTYPE_F32)
    unipc_step_kernel<float, float><<<blocks, threads, 0, s>>>(*a);
  else if (a->sample_dtype == CE_DTYPE_F32 && a->model_dtype == CE_DTYPE_BF16)
    unipc_step_kernel<float, __nv_bfloat16><<<blocks, threads, 0, s>>>(*a);
  else if (a->sample_dtype == CE_DTYPE_BF16 && a->model_dtype == CE_DTYPE_BF16)
    unipc_step_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, threads, 0, s>>>(*a);
  else
    return fail(CE_ERR_INVALID, "unipc_step: supported (sample, model) dtypes are (f32,f32), (f32,bf16), (bf16,bf16)");
  CE_CHECK_CUDA(cudaGetLastError());
  return CE_OK;
}
