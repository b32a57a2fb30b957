// FastVLA action head as one kernel (fastvla/fastvlm_with_expert.py:23-38, 50-54):
//   s = SiLU(Linear(LayerNorm(state)))                     state_projection
//   f = cat(pooled, s)
//   f = SiLU(Linear(Dropout(SiLU(LayerNorm(Linear(f))))))  fusion   (Dropout is identity in eval)
//   a = Linear(f)                                          action_head
// A cluster of 8 CTAs carries R batch rows through all four matrices, each CTA owning an eighth of every layer's
// output neurons; each warp streams weight rows as 128-bit vectors and reduces R dot products at once.
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

// One kernel, one thread-block CLUSTER of 8 CTAs per R batch rows: every matrix is split 8 ways over its output
// neurons, so 8 SMs stream the 6 MB of head weights for a row group instead of one (at b = 1 the single-CTA version
// took 0.37 ms = 7 % of the whole select_action; the weights come from L2 at a few bytes per clock per SM, so the
// only lever is more SMs per row).  Layer outputs are exchanged through small global scratch rows (they double as
// the parity taps) ordered by cluster-scope release/acquire barriers; LayerNorm statistics are recomputed by every
// CTA from the full row (1024 values).
constexpr int HEAD_CL = 8;

template <typename T, int R>
__global__ void __cluster_dims__(HEAD_CL, 1, 1) __launch_bounds__(256)
action_head_kernel(HeadWeights w, const float* __restrict__ pooled, const float* __restrict__ states,
                   float* __restrict__ actions, float* state_feat, float* x1_scratch, float* fused, int B) {
  extern __shared__ __align__(16) float sh[];
  const int H = w.H, S = w.S, Hd = w.Hd, F = w.F, A = w.A;
  const int KC = H + Hd;
  float* cat = sh;              // [R][KC]
  float* x1 = cat + R * KC;     // [R][F]
  float* x2 = x1 + R * F;       // [R][F]
  float* sln = x2 + R * F;      // [R][S] normalised state
  const int crank = static_cast<int>(blockIdx.x % HEAD_CL);
  const int r0 = static_cast<int>(blockIdx.x / HEAD_CL) * R;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nwarps = blockDim.x >> 5;
  constexpr int NC = 8;   // weight rows in flight per warp
  auto cluster_sync = [] {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  };

  // ---- under the previous kernel's tail: this CTA's slices of the two big matrices (constants) -> L2, so the
  // latency-bound row streams below hit L2 instead of paying a DRAM round trip per step ----
  {
    const int fper0 = ((F + HEAD_CL - 1) / HEAD_CL + NC - 1) / NC * NC;
    const int lo = crank * fper0, hi = min(F, lo + fper0);
    if (hi > lo) {
      const char* m0 = reinterpret_cast<const char*>(static_cast<const T*>(w.w_f0) + static_cast<size_t>(lo) * KC);
      const char* m4 = reinterpret_cast<const char*>(static_cast<const T*>(w.w_f4) + static_cast<size_t>(lo) * F);
      const size_t b0 = static_cast<size_t>(hi - lo) * KC * sizeof(T), b4 = static_cast<size_t>(hi - lo) * F * sizeof(T);
      constexpr size_t CH = 4096;
      const size_t n0c = (b0 + CH - 1) / CH, n4c = (b4 + CH - 1) / CH;
      if (((reinterpret_cast<uintptr_t>(m0) | reinterpret_cast<uintptr_t>(m4) | b0 | b4) & 15u) == 0) {
        for (size_t i = tid; i < n0c + n4c; i += blockDim.x) {
          const bool second = i >= n0c;
          const size_t off = (second ? i - n0c : i) * CH, tot = second ? b4 : b0;
          const char* ptr = (second ? m4 : m0) + off;
          const uint32_t bytes = static_cast<uint32_t>(tot - off < CH ? tot - off : CH);
          asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
        }
      }
    }
  }
  pdl_sync();
  // ---- LayerNorm(state) (eps 1e-5, biased variance); every CTA of the cluster needs all of it ----
  if (warp < R) {
    const int b = r0 + warp;
    if (b < B) {
      // dataset normalisation of the raw state (identity unless fvla_set_io_normalization was called)
      auto st = [&](int i) { return (states[static_cast<size_t>(b) * S + i] - w.st_mean[i]) * w.st_inv_std[i]; };
      float s = 0.f;
      for (int i = lane; i < S; i += 32) s += st(i);
      const float mean = warp_sum(s) / static_cast<float>(S);
      float vs = 0.f;
      for (int i = lane; i < S; i += 32) {
        const float d = st(i) - mean;
        vs = fmaf(d, d, vs);
      }
      const float rstd = rsqrtf(warp_sum(vs) / static_cast<float>(S) + 1e-5f);
      for (int i = lane; i < S; i += 32)
        sln[warp * S + i] = (st(i) - mean) * rstd * w.ln_s_w[i] + w.ln_s_b[i];
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
  // ---- state projection + SiLU: this CTA's slice of the Hd outputs -> state_feat (global) ----
  {
    const T* ws = static_cast<const T*>(w.w_state);
    const int per = (Hd + HEAD_CL - 1) / HEAD_CL;
    const int n_end = min(Hd, (crank + 1) * per);
    for (int n = crank * per + tid; n < n_end; n += blockDim.x) {
      float acc[R];
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = w.b_state[n];
      for (int k = 0; k < S; ++k) {
        const float wv = ld_w(ws + static_cast<size_t>(n) * S + k);
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r] = fmaf(wv, sln[r * S + k], acc[r]);
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (r0 + r < B) state_feat[static_cast<size_t>(r0 + r) * Hd + n] = silu_precise(acc[r]);
    }
  }
  cluster_sync();
  for (int i = tid; i < R * Hd; i += blockDim.x) {
    const int r = i / Hd, c = i % Hd;
    cat[r * KC + H + c] = r0 + r < B ? state_feat[static_cast<size_t>(r0 + r) * Hd + c] : 0.f;
  }
  __syncthreads();
  // ---- fusion.0: Linear(KC -> F), this CTA's slice -> x1_scratch (global) ----
  const int fper = ((F + HEAD_CL - 1) / HEAD_CL + NC - 1) / NC * NC;  // slice width, a multiple of NC
  const int f_lo = crank * fper, f_hi = min(F, f_lo + fper);
  {
    const T* wf = static_cast<const T*>(w.w_f0);
    for (int n0 = f_lo + warp * NC; n0 < f_hi; n0 += nwarps * NC) {
      float acc[NC][R];
      const int nv = f_hi - n0 < NC ? f_hi - n0 : NC;
      warp_dot_rows<T, R, NC>(wf + static_cast<size_t>(n0) * KC, KC, nv, cat, KC, KC, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j)
          if (j < nv) {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (r0 + r < B) x1_scratch[static_cast<size_t>(r0 + r) * F + n0 + j] = acc[j][r] + w.b_f0[n0 + j];
          }
      }
    }
  }
  cluster_sync();
  for (int i = tid; i < R * F; i += blockDim.x) {
    const int r = i / F, c = i % F;
    x1[r * F + c] = r0 + r < B ? x1_scratch[static_cast<size_t>(r0 + r) * F + c] : 0.f;
  }
  __syncthreads();
  // ---- fusion.1 LayerNorm + fusion.2 SiLU (in place, every CTA on the full rows) ----
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
  // ---- fusion.4: Linear(F -> F) + SiLU, this CTA's slice -> fused (global) ----
  {
    const T* wf = static_cast<const T*>(w.w_f4);
    for (int n0 = f_lo + warp * NC; n0 < f_hi; n0 += nwarps * NC) {
      float acc[NC][R];
      const int nv = f_hi - n0 < NC ? f_hi - n0 : NC;
      warp_dot_rows<T, R, NC>(wf + static_cast<size_t>(n0) * F, F, nv, x1, F, F, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < NC; ++j)
          if (j < nv) {
#pragma unroll
            for (int r = 0; r < R; ++r)
              if (r0 + r < B) fused[static_cast<size_t>(r0 + r) * F + n0 + j] = silu_precise(acc[j][r] + w.b_f4[n0 + j]);
          }
      }
    }
  }
  cluster_sync();
  for (int i = tid; i < R * F; i += blockDim.x) {
    const int r = i / F, c = i % F;
    x2[r * F + c] = r0 + r < B ? fused[static_cast<size_t>(r0 + r) * F + c] : 0.f;
  }
  __syncthreads();
  // ---- action_head: Linear(F -> A), output neurons dealt round-robin over the cluster's warps ----
  {
    const T* wa = static_cast<const T*>(w.w_act);
    for (int n = crank * nwarps + warp; n < A; n += HEAD_CL * nwarps) {
      float acc[1][R];
      warp_dot_rows<T, R, 1>(wa + static_cast<size_t>(n) * F, F, 1, x2, F, F, lane, acc);
      if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (r0 + r < B)
            actions[static_cast<size_t>(r0 + r) * A + n] = fmaf(acc[0][r] + w.b_act[n], w.act_scale[n], w.act_shift[n]);
      }
    }
  }
}

template <typename T>
int launch_head(const HeadWeights& w, const float* pooled, const float* states, float* actions,
                float* state_feat, float* x1_scratch, float* fused, int B, cudaStream_t stream) {
  constexpr int R = 4;
  auto kfn = action_head_kernel<T, R>;
  const size_t smem = sizeof(float) * (static_cast<size_t>(R) * (w.H + w.Hd + 2 * w.F + w.S));
  FVLA_REQUIRE(smem <= 220 * 1024, "action head: hidden sizes too large for one CTA");
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), static_cast<int>(smem))) return rc;
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(HEAD_CL * ceil_div(B, R)), dim3(256), smem, stream, w, pooled, states, actions,
                             state_feat, x1_scratch, fused, B));
  return 0;
}

}  // namespace

int action_head(int dtype, const HeadWeights& w, const float* pooled, const float* states,
                float* actions, float* state_feat, float* x1_scratch, float* fused, int B,
                cudaStream_t stream) {
  FVLA_REQUIRE(B > 0, "action head: empty batch");
  FVLA_REQUIRE(state_feat != nullptr && x1_scratch != nullptr && fused != nullptr,
               "action head: the three [B, *] fp32 exchange rows are required");
  FVLA_REQUIRE((w.H + w.Hd) % 8 == 0 && w.F % 8 == 0 && w.H % 4 == 0,
               "action head: H+hidden_dim and fusion_dim must be multiples of 8");
  if (dtype == DT_F32)
    return launch_head<float>(w, pooled, states, actions, state_feat, x1_scratch, fused, B, stream);
  return launch_head<__nv_bfloat16>(w, pooled, states, actions, state_feat, x1_scratch, fused, B, stream);
}

}  // namespace fvla
