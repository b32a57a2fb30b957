// Shared device/host helpers for the FastVLA sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <string>

namespace fvla {

// ---- status / error plumbing (C-ABI returns int; message kept per thread) ----
void set_error(const std::string& msg);
const char* last_error();

#define FVLA_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::fvla::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " at " +      \
                        __FILE__ + ":" + std::to_string(__LINE__));                        \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

#define FVLA_REQUIRE(cond, msg)                                                            \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      ::fvla::set_error(std::string("requirement failed: ") + #cond + " — " + (msg));      \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

// ACT_GELU_HALF: the operand is x/2 (weights and bias pre-scaled by 1/2 at pack time), result gelu(x)
// ACT_GELU_HALF_F16 (bf16 tensor-core GEMM only): as ACT_GELU_HALF, but the result is stored as FP16 — for hidden
// tensors whose only consumer is another of our GEMMs run with fp16 operands (GemmArgs::ab_f16)
enum Act : int { ACT_NONE = 0, ACT_GELU = 1, ACT_SILU = 2, ACT_RELU = 3, ACT_GELU_HALF = 4, ACT_GELU_HALF_F16 = 5 };

// ---- scalar conversions -------------------------------------------------------
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) {
  return __float2bfloat16_rn(v);
}

// erf-based GELU (torch nn.GELU default, approximate='none')
__device__ __forceinline__ float gelu_erf(float x) {
  return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// Output-side GELU of the NHWC kernels: exact erf form when activations are stored in fp32 (parity
// mode); for bf16 storage the tanh form with the polynomial fitted to erf (|err| <= 2.5e-5 + 2^-11 rel,
// see gemm_sm100.cu) — a quarter of the instructions of erff.
__device__ __forceinline__ float gelu_tanh_fit(float x) {
  const float u = fminf(x * x, 64.0f);
  const float q = fmaf(u, fmaf(u, -3.51516789e-04f, 3.70056460e-02f), 7.97507884e-01f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * q));
  const float h = 0.5f * x;
  return fmaf(h, t, h);
}
template <typename T> __device__ __forceinline__ float gelu_store(float x);
template <> __device__ __forceinline__ float gelu_store<float>(float x) { return gelu_erf(x); }
template <> __device__ __forceinline__ float gelu_store<__nv_bfloat16>(float x) { return gelu_tanh_fit(x); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_precise(float x) { return x / (1.0f + expf(-x)); }

template <int ACT> __device__ __forceinline__ float apply_act(float x) {
  if constexpr (ACT == ACT_GELU) return gelu_erf(x);
  else if constexpr (ACT == ACT_SILU) return silu_precise(x);
  else if constexpr (ACT == ACT_RELU) return fmaxf(x, 0.0f);
  else return x;
}
__device__ __forceinline__ float apply_act_rt(float x, int act) {
  switch (act) {
    case ACT_GELU: return gelu_erf(x);
    case ACT_SILU: return silu_precise(x);
    case ACT_RELU: return fmaxf(x, 0.0f);
    case ACT_GELU_HALF: return gelu_erf(2.0f * x);
    default: return x;
  }
}

// ---- 8-wide channel vectors (16 B of bf16 / 32 B of fp32) ------------------------
// All NHWC kernels move 8 consecutive channels per thread so bf16 traffic is 128-bit.
template <typename T> struct Vec8;
template <> struct Vec8<float> {
  float v[8];
  __device__ __forceinline__ void load(const float* p) {
    float4 a = __ldg(reinterpret_cast<const float4*>(p));
    float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  __device__ __forceinline__ void store(float* p) const {
    reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
};
template <> struct Vec8<__nv_bfloat16> {
  float v[8];
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    uint4 raw = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
  __device__ __forceinline__ void store(__nv_bfloat16* p) const {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` must hold >= 32 floats. All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? red[lane] : 0.0f;
  r = warp_sum(r);
  return r;
}

// ---- programmatic dependent launch ---------------------------------------------------------------------------
// Every hot-path kernel is launched with the programmatic-stream-serialisation attribute (launch_pdl below): its
// CTAs may become resident while the previous kernel of the stream drains, run their input-independent prologue
// (barrier init, tensor-map prefetch, TMEM allocation, constant tables) and then block in pdl_sync() until the
// previous grid has completed and its writes are visible.  ALL threads of a CTA execute pdl_sync() before the first
// access to memory another kernel of the forward may have written, and before the first global write.  The trigger
// follows the wait, so at most two grids are in flight and a CTA that triggers has already allocated its TMEM
// (a dependent CTA can never take columns a still-unallocated primary CTA of the same SM waits for).
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

int num_sms();  // cached SM count of the current device
// cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device): raised once per pair (a process may hold
// engines on several GPUs), thread-safe, ~50 ns on the repeat path
int ensure_dyn_smem(const void* fn, int bytes);
bool pdl_enabled();  // FVLA_DISABLE_PDL unset (A/B switch)
// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-dependent-launch attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}


}  // namespace fvla
