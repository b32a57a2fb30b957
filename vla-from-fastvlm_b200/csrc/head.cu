// FastVLA action head as one kernel (fastvla/fastvlm_with_expert.py:23-38, 50-54):
//   s = SiLU(Linear(LayerNorm(state)))                     state_projection
//   f = cat(pooled, s)
//   f = SiLU(Linear(Dropout(SiLU(LayerNorm(Linear(f))))))  fusion   (Dropout is identity in eval)
//   a = Linear(f)                                          action_head
// A CTA carries R batch rows through all four matrices; activations stay in shared memory, each warp
// streams weight rows as 128-bit vectors and reduces R dot products at once.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

template <typename T> __device__ __forceinline__ float ld_w(const T* p) { return to_f32(*p); }

// y[r][n0+j] = sum_k x[r][k] * W[n0+j][k] for NC consecutive output columns and all R rows; K % 8 == 0.
// NC independent weight streams per warp keep several 128-bit loads in flight (the head is
// latency-bound: 6 MB of weights read once by a handful of CTAs).
template <typename T, int R, int NC>
__device__ __forceinline__ void warp_dot_rows(const T* __restrict__ wbase, int ldw, int n_valid,
                                              const float* x, int ldx, int K, int lane,
                                              float (&acc)[NC][R]) {
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int r = 0; r < R; ++r) acc[j][r] = 0.f;
  for (int k = lane * 8; k < K; k += 32 * 8) {
    Vec8<T> w[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (j < n_valid) w[j].load(wbase + static_cast<size_t>(j) * ldw + k);
      else {
#pragma unroll
        for (int c = 0; c < 8; ++c) w[j].v[c] = 0.f;
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 x0 = *reinterpret_cast<const float4*>(x + r * ldx + k);
      const float4 x1 = *reinterpret_cast<const float4*>(x + r * ldx + k + 4);
#pragma unroll
      for (int j = 0; j < NC; ++j) {
        float a = acc[j][r];
        a = fmaf(w[j].v[0], x0.x, a); a = fmaf(w[j].v[1], x0.y, a); a = fmaf(w[j].v[2], x0.z, a); a = fmaf(w[j].v[3], x0.w, a);
        a = fmaf(w[j].v[4], x1.x, a); a = fmaf(w[j].v[5], x1.y, a); a = fmaf(w[j].v[6], x1.z, a); a = fmaf(w[j].v[7], x1.w, a);
        acc[j][r] = a;
      }
    }
  }
#pragma unroll
  for (int j = 0; j < NC; ++j)
#pragma unroll
    for (int r = 0; r < R; ++r) acc[j][r] = warp_sum(acc[j][r]);
}

template <typename T, int R>
__global__ void __launch_bounds__(256)
action_head_kernel(HeadWeights w, const float* __restrict__ pooled, const float* __restrict__ states,
                   float* __restrict__ actions, float* tap_state, float* tap_fused, int B) {
  extern __shared__ __align__(16) float sh[];
  const int H = w.H, S = w.S, Hd = w.Hd, F = w.F, A = w.A;
  const int KC = H + Hd;
  float* cat = sh;              // [R][KC]
  float* x1 = cat + R * KC;     // [R][F]
  float* x2 = x1 + R * F;       // [R][F]
  float* sln = x2 + R * F;      // [R][S] normalised state
  const int r0 = blockIdx.x * R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  constexpr int NC = 4;

  // ---- LayerNorm(state) (eps 1e-5, biased variance) ----
  if (warp < R) {
    const int b = r0 + warp;
    if (b < B) {
      float s = 0.f;
      for (int i = lane; i < S; i += 32) s += states[static_cast<size_t>(b) * S + i];
      const float mean = warp_sum(s) / static_cast<float>(S);
      float vs = 0.f;
      for (int i = lane; i < S; i += 32) {
        const float d = states[static_cast<size_t>(b) * S + i] - mean;
        vs = fmaf(d, d, vs);
      }
      const float rstd = rsqrtf(warp_sum(vs) / static_cast<float>(S) + 1e-5f);
      for (int i = lane; i < S; i += 32)
        sln[warp * S + i] = (states[static_cast<size_t>(b) * S + i] - mean) * rstd * w.ln_s_w[i] + w.ln_s_b[i];
    } else {
      for (int i = lane; i < S; i += 32) sln[warp * S + i] = 0.f;
    }
  }
  // ---- pooled -> cat[:, :H] ----
  for (int i = tid; i < R * H; i += blockDim.x) {
    const int r = i / H, c = i % H;
    const int b = r0 + r;
    cat[r * KC + c] = b < B ? pooled[static_cast<size_t>(b) * H + c] : 0.f;
  }
  __syncthreads();
  // ---- state projection + SiLU -> cat[:, H:] ----
  {
    const T* ws = static_cast<const T*>(w.w_state);
    for (int n = tid; n < Hd; n += blockDim.x) {
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = w.b_state[n];
      for (int k = 0; k < S; ++k) {
        const float wv = ld_w(ws + static_cast<size_t>(n) * S + k);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(wv, sln[r * S + k], acc[r]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float y = silu_precise(acc[r]);
        cat[r * KC + H + n] = y;
        if (tap_state != nullptr && r0 + r < B) tap_state[static_cast<size_t>(r0 + r) * Hd + n] = y;
      }
    }
  }
  __syncthreads();
  // ---- fusion.0: Linear(KC -> F) ----
  {
    const T* wf = static_cast<const T*>(w.w_f0);
    for (int n0 = warp * NC; n0 < F; n0 += nwarps * NC) {
      float acc[NC][R];
      const int nv = F - n0 < NC ? F - n0 : NC;
      warp_dot_rows<T, R, NC>(wf + static_cast<size_t>(n0) * KC, KC, nv, cat, KC, KC, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j)
          if (j < nv) {
#pragma unroll
            for (int r = 0; r < R; ++r) x1[r * F + n0 + j] = acc[j][r] + w.b_f0[n0 + j];
          }
      }
    }
  }
  __syncthreads();
  // ---- fusion.1 LayerNorm + fusion.2 SiLU (in place) ----
  for (int r = warp; r < R; r += nwarps) {
    float s = 0.f;
    for (int i = lane; i < F; i += 32) s += x1[r * F + i];
    const float mean = warp_sum(s) / static_cast<float>(F);
    float vs = 0.f;
    for (int i = lane; i < F; i += 32) { const float d = x1[r * F + i] - mean; vs = fmaf(d, d, vs); }
    const float rstd = rsqrtf(warp_sum(vs) / static_cast<float>(F) + 1e-5f);
    for (int i = lane; i < F; i += 32)
      x1[r * F + i] = silu_precise((x1[r * F + i] - mean) * rstd * w.ln_f_w[i] + w.ln_f_b[i]);
  }
  __syncthreads();
  // ---- fusion.4: Linear(F -> F) + SiLU ----
  {
    const T* wf = static_cast<const T*>(w.w_f4);
    for (int n0 = warp * NC; n0 < F; n0 += nwarps * NC) {
      float acc[NC][R];
      const int nv = F - n0 < NC ? F - n0 : NC;
      warp_dot_rows<T, R, NC>(wf + static_cast<size_t>(n0) * F, F, nv, x1, F, F, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j)
          if (j < nv) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float y = silu_precise(acc[j][r] + w.b_f4[n0 + j]);
              x2[r * F + n0 + j] = y;
              if (tap_fused != nullptr && r0 + r < B) tap_fused[static_cast<size_t>(r0 + r) * F + n0 + j] = y;
            }
          }
      }
    }
  }
  __syncthreads();
  // ---- action_head: Linear(F -> A) ----
  {
    const T* wa = static_cast<const T*>(w.w_act);
    for (int n = warp; n < A; n += nwarps) {
      float acc[1][R];
      warp_dot_rows<T, R, 1>(wa + static_cast<size_t>(n) * F, F, 1, x2, F, F, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r0 + r < B) actions[static_cast<size_t>(r0 + r) * A + n] = acc[0][r] + w.b_act[n];
      }
    }
  }
}

template <typename T>
int launch_head(const HeadWeights& w, const float* pooled, const float* states, float* actions,
                float* tap_state, float* tap_fused, int B, cudaStream_t stream) {
  constexpr int R = 2;
  auto kfn = action_head_kernel<T, R>;
  const size_t smem = sizeof(float) * (static_cast<size_t>(R) * (w.H + w.Hd + 2 * w.F + w.S));
  FVLA_REQUIRE(smem <= 220 * 1024, "action head: hidden sizes too large for one CTA");
  FVLA_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
  kfn<<<ceil_div(B, R), 256, smem, stream>>>(w, pooled, states, actions, tap_state, tap_fused, B);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

int action_head(int dtype, const HeadWeights& w, const float* pooled, const float* states,
                float* actions, float* tap_state_feat, float* tap_fused, int B,
                cudaStream_t stream) {
  FVLA_REQUIRE(B > 0, "action head: empty batch");
  FVLA_REQUIRE((w.H + w.Hd) % 8 == 0 && w.F % 8 == 0 && w.H % 4 == 0,
               "action head: H+hidden_dim and fusion_dim must be multiples of 8");
  if (dtype == DT_F32)
    return launch_head<float>(w, pooled, states, actions, tap_state_feat, tap_fused, B, stream);
  return launch_head<__nv_bfloat16>(w, pooled, states, actions, tap_state_feat, tap_fused, B, stream);
}

}  // namespace fvla
