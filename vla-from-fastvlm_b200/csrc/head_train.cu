// Training step of the FastVLA action head: forward + MSE loss + backward in fp32, gradients written straight into
// the caller's flat gradient buffer (the one the data-parallel all-reduce runs over).
//
// Reference: the head modules of fastvla/fastvlm_with_expert.py:23-38,50-54 in train mode (Dropout p active), the
// loss of lerobot_fastvla/modeling_fastvla.py:127-133 (`F.mse_loss(pred, gt[:, 0])`, mean over B*A) and
// `accelerator.backward(loss)` in training/trainer.py:175.  The frozen backbone contributes the pooled features only
// (it runs under no_grad in the reference too), so the whole autograd graph of the step is this head:
//
//   sn  = LN(state; g_s, b_s)            z1 = W1 sn + b1          s  = silu(z1)
//   y0  = Wf0 [pooled | s] + bf0         yn = LN(y0; g_f, b_f)    x1 = silu(yn) * keep / (1 - p)
//   y4  = Wf4 x1 + bf4                   x2 = silu(y4)            a  = Wa x2 + ba
//   loss = mean((a - target)^2)
//
// The batch is small (16-64 rows per GPU) and the matrices hold 3 M parameters: every linear layer is a "skinny"
// product, so the kernels are organised to stream each weight matrix exactly once per pass with all batch rows held
// on chip (forward: a warp per output neuron; data gradient: a thread per input column; weight gradient: a thread per
// weight element).  ~25 launches, ~40 MB of L2 traffic, no atomics on gradients: results are deterministic.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

constexpr int MAXB = 64;  // rows per pass of the skinny kernels (larger batches loop)
constexpr int kDgradSplits = 8;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ float silu_f(float x) { return x * sigmoidf_(x); }
__device__ __forceinline__ float silu_grad(float x) {
  const float s = sigmoidf_(x);
  return s * (1.0f + x * (1.0f - s));
}

// Y[b][n] = X[b][:] . W[n][:] + bias[n]   (nn.Linear forward), warp per output neuron, RB batch rows per pass
// (W is re-streamed once per pass: 16 rows for the larger batches).
// K % 4 == 0 and 16-byte aligned rows take the float4 path (one 128-bit load of W serves 8 x 4 FMAs).
template <int RB>
__global__ void __launch_bounds__(256)
skinny_nt_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ W, const float* __restrict__ bias,
                 float* __restrict__ Y, int ldy, int B, int N, int K) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  const float* wr = W + static_cast<size_t>(n) * K;
  const bool vec = (K & 3) == 0 && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(X) & 15u) == 0 &&
                   (reinterpret_cast<uintptr_t>(W) & 15u) == 0;
  for (int b0 = 0; b0 < B; b0 += RB) {
    float acc[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) acc[i] = 0.f;
    if (vec) {
      for (int k = lane * 4; k < K; k += 128) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(wr + k));
#pragma unroll
        for (int i = 0; i < RB; ++i)
          if (b0 + i < B) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(X + static_cast<size_t>(b0 + i) * ldx + k));
            acc[i] = fmaf(w.x, x.x, fmaf(w.y, x.y, fmaf(w.z, x.z, fmaf(w.w, x.w, acc[i]))));
          }
      }
    } else {
      for (int k = lane; k < K; k += 32) {
        const float w = __ldg(wr + k);
#pragma unroll
        for (int i = 0; i < RB; ++i)
          if (b0 + i < B) acc[i] = fmaf(w, __ldg(X + static_cast<size_t>(b0 + i) * ldx + k), acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const float v = warp_sum(acc[i]);
      if (lane == 0 && b0 + i < B) Y[static_cast<size_t>(b0 + i) * ldy + n] = v + (bias ? bias[n] : 0.f);
    }
  }
}

// dX[b][k] = sum_n dY[b][n] W[n][k] for k in [k0, k0 + Kc)   (data gradient of nn.Linear), thread per column k
// The n loop is split over gridDim.z CTAs (a column block alone would occupy a dozen SMs): split z writes its partial
// sums to dX + z * split_stride and skinny_reduce_kernel adds them up in a fixed order (deterministic).
template <int BT>
__global__ void __launch_bounds__(128)
skinny_nn_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ W, int ldw, float* __restrict__ dX,
                 int ldx, int B, int N, int k0, int Kc, int n_per_split, size_t split_stride) {
  const int k = blockIdx.x * 128 + threadIdx.x;
  const int b0 = blockIdx.y * BT;
  const int n_lo = blockIdx.z * n_per_split, n_hi = min(N, n_lo + n_per_split);
  dX += static_cast<size_t>(blockIdx.z) * split_stride;
  __shared__ float sdy[BT][64];
  float acc[BT];
#pragma unroll
  for (int i = 0; i < BT; ++i) acc[i] = 0.f;
  for (int n0 = n_lo; n0 < n_hi; n0 += 64) {
    for (int idx = threadIdx.x; idx < BT * 64; idx += 128) {
      const int i = idx >> 6, j = idx & 63;
      sdy[i][j] = (b0 + i < B && n0 + j < n_hi) ? dY[static_cast<size_t>(b0 + i) * ldy + n0 + j] : 0.f;
    }
    __syncthreads();
    if (k < Kc) {
      const int nmax = min(64, n_hi - n0);
      for (int j = 0; j < nmax; ++j) {
        const float w = __ldg(W + static_cast<size_t>(n0 + j) * ldw + k0 + k);
#pragma unroll
        for (int i = 0; i < BT; ++i) acc[i] = fmaf(sdy[i][j], w, acc[i]);
      }
    }
    __syncthreads();
  }
  if (k < Kc)
#pragma unroll
    for (int i = 0; i < BT; ++i)
      if (b0 + i < B) dX[static_cast<size_t>(b0 + i) * ldx + k] = acc[i];
}

__global__ void skinny_reduce_kernel(const float* __restrict__ part, size_t split_stride, int splits,
                                     float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int z = 0; z < splits; ++z) s += part[static_cast<size_t>(z) * split_stride + i];
  out[i] = s;
}

// dW[n][k] = sum_b dY[b][n] X[b][k]; db[n] = sum_b dY[b][n]   (weight / bias gradient), thread per weight element
__global__ void __launch_bounds__(256)
skinny_tn_kernel(const float* __restrict__ dY, int ldy, const float* __restrict__ X, int ldx, float* __restrict__ dW,
                 float* __restrict__ db, int B, int N, int K) {
  const int k = blockIdx.x * 256 + threadIdx.x;
  const int n = blockIdx.y;
  __shared__ float sdy[MAXB];
  float acc = 0.f, accb = 0.f;
  for (int b0 = 0; b0 < B; b0 += MAXB) {
    __syncthreads();
    if (threadIdx.x < MAXB) sdy[threadIdx.x] = b0 + threadIdx.x < B ? dY[static_cast<size_t>(b0 + threadIdx.x) * ldy + n] : 0.f;
    __syncthreads();
    const int bm = min(MAXB, B - b0);
    if (k < K)
      for (int b = 0; b < bm; ++b) acc = fmaf(sdy[b], __ldg(X + static_cast<size_t>(b0 + b) * ldx + k), acc);
    if (k == 0)
      for (int b = 0; b < bm; ++b) accb += sdy[b];
  }
  if (k < K) dW[static_cast<size_t>(n) * K + k] = acc;
  if (k == 0 && db != nullptr) db[n] = accb;
}

// LayerNorm forward (eps 1e-5, biased variance): xhat, rstd and y = xhat * g + b; warp per row
__global__ void __launch_bounds__(128)
ln_fwd_kernel(const float* __restrict__ x, int ldx, const float* __restrict__ g, const float* __restrict__ bta,
              float* __restrict__ xhat, float* __restrict__ rstd, float* __restrict__ y, int B, int N) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const float* xr = x + static_cast<size_t>(row) * ldx;
  float s = 0.f;
  for (int i = lane; i < N; i += 32) s += xr[i];
  const float mean = warp_sum(s) / static_cast<float>(N);
  float vs = 0.f;
  for (int i = lane; i < N; i += 32) { const float d = xr[i] - mean; vs = fmaf(d, d, vs); }
  const float r = rsqrtf(warp_sum(vs) / static_cast<float>(N) + 1e-5f);
  if (lane == 0) rstd[row] = r;
  for (int i = lane; i < N; i += 32) {
    const float h = (xr[i] - mean) * r;
    xhat[static_cast<size_t>(row) * N + i] = h;
    y[static_cast<size_t>(row) * N + i] = fmaf(h, g[i], bta[i]);
  }
}

// LayerNorm backward for the input: dx = rstd * (dxh - mean(dxh) - xhat * mean(dxh * xhat)), dxh = dy * g
__global__ void __launch_bounds__(128)
ln_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xhat, const float* __restrict__ rstd,
              const float* __restrict__ g, float* __restrict__ dx, int B, int N) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const size_t o = static_cast<size_t>(row) * N;
  float s1 = 0.f, s2 = 0.f;
  for (int i = lane; i < N; i += 32) {
    const float d = dy[o + i] * g[i];
    s1 += d;
    s2 = fmaf(d, xhat[o + i], s2);
  }
  s1 = warp_sum(s1) / static_cast<float>(N);
  s2 = warp_sum(s2) / static_cast<float>(N);
  const float r = rstd[row];
  for (int i = lane; i < N; i += 32) dx[o + i] = r * (dy[o + i] * g[i] - s1 - xhat[o + i] * s2);
}

// column sums over the batch: out_a[n] = sum_b a[b][n] * (m ? m[b][n] : 1), out_b[n] = sum_b a[b][n]
__global__ void colsum2_kernel(const float* __restrict__ a, const float* __restrict__ m, float* __restrict__ out_am,
                               float* __restrict__ out_a, int B, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float s = 0.f, sm = 0.f;
  for (int b = 0; b < B; ++b) {
    const float v = a[static_cast<size_t>(b) * N + n];
    s += v;
    if (m != nullptr) sm = fmaf(v, m[static_cast<size_t>(b) * N + n], sm);
  }
  if (out_a != nullptr) out_a[n] = s;
  if (out_am != nullptr) out_am[n] = sm;
}

// cat[b] = [pooled[b] | silu(z1[b])]
__global__ void silu_cat_kernel(const float* __restrict__ pooled, const float* __restrict__ z1, float* __restrict__ cat,
                                int B, int H, int Hd) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int KC = H + Hd;
  if (idx >= B * KC) return;
  const int b = idx / KC, c = idx % KC;
  cat[idx] = c < H ? pooled[static_cast<size_t>(b) * H + c] : silu_f(z1[static_cast<size_t>(b) * Hd + c - H]);
}
// out = silu(x) * (keep ? keep / (1-p) : 1)
__global__ void silu_drop_kernel(const float* __restrict__ x, const uint8_t* __restrict__ keep, float inv_keep,
                                 float* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = silu_f(x[i]);
  out[i] = keep != nullptr ? (keep[i] ? s * inv_keep : 0.f) : s;
}
// dx = dy * silu'(x) * (keep ? keep / (1-p) : 1), reading dy with a row pitch / column offset
__global__ void silu_bwd_kernel(const float* __restrict__ dy, int ldy, const float* __restrict__ x,
                                const uint8_t* __restrict__ keep, float inv_keep, float* __restrict__ dx, int B, int N) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * N) return;
  const int b = i / N, c = i % N;
  float d = dy[static_cast<size_t>(b) * ldy + c];
  if (keep != nullptr) d = keep[i] ? d * inv_keep : 0.f;
  dx[i] = d * silu_grad(x[i]);
}
// loss = mean((a - t)^2), da = 2 (a - t) / (B A); one block
__global__ void __launch_bounds__(256)
mse_kernel(const float* __restrict__ a, const float* __restrict__ t, float* __restrict__ da, float* __restrict__ loss,
           int n) {
  __shared__ float red[32];
  float s = 0.f;
  const float inv = 1.0f / static_cast<float>(n);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float d = a[i] - t[i];
    s = fmaf(d, d, s);
    da[i] = 2.0f * d * inv;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) *loss = s * inv;
}

}  // namespace

size_t head_train_scratch_floats(int B, int H, int S, int Hd, int F, int A) {
  const size_t b = static_cast<size_t>(B);
  const size_t widest = static_cast<size_t>(H + Hd > F ? H + Hd : F);
  return b * (3 * S + 1 + 2 * Hd + (H + Hd) + 6 * F + 1 + 2 * A + (H + Hd)) + kDgradSplits * b * widest + 128;
}

int head_train_step(const HeadTrainArgs& t, cudaStream_t s) {
  const int B = t.B, H = t.H, S = t.S, Hd = t.Hd, F = t.F, A = t.A, KC = H + Hd;
  FVLA_REQUIRE(B > 0 && H > 0 && S > 0 && Hd > 0 && F > 0 && A > 0, "head_train_step: empty dimension");
  FVLA_REQUIRE(t.scratch != nullptr && t.scratch_floats >= head_train_scratch_floats(B, H, S, Hd, F, A),
               "head_train_step: scratch too small (see head_train_scratch_floats)");
  FVLA_REQUIRE(t.drop_p >= 0.f && t.drop_p < 1.f, "head_train_step: dropout probability must be in [0, 1)");
  FVLA_REQUIRE((reinterpret_cast<uintptr_t>(t.scratch) & 15u) == 0, "head_train_step: scratch must be 16-byte aligned");
  const size_t b = static_cast<size_t>(B);
  float* p = t.scratch;
  auto take = [&](size_t n) { float* r = p; p += (n + 3) & ~static_cast<size_t>(3); return r; };  // 16-byte aligned rows
  float *shat = take(b * S), *sn = take(b * S), *rstd_s = take(b), *z1 = take(b * Hd), *cat = take(b * KC);
  float *y0 = take(b * F), *yhat = take(b * F), *rstd_f = take(b), *yn = take(b * F), *x1 = take(b * F);
  float *y4 = take(b * F), *x2 = take(b * F), *a = take(b * A), *da = take(b * A);
  float *dtmp = take(b * KC), *dz1 = take(b * Hd), *dsn = take(b * S);
  float* dpart = take(static_cast<size_t>(kDgradSplits) * b * static_cast<size_t>(KC > F ? KC : F));
  // gradient slots in nn.Module parameter order
  float* g = t.grads;
  float *g_lnsw = g, *g_lnsb = g_lnsw + S, *g_w1 = g_lnsb + S, *g_b1 = g_w1 + static_cast<size_t>(Hd) * S;
  float *g_wf0 = g_b1 + Hd, *g_bf0 = g_wf0 + static_cast<size_t>(F) * KC, *g_lnfw = g_bf0 + F, *g_lnfb = g_lnfw + F;
  float *g_wf4 = g_lnfb + F, *g_bf4 = g_wf4 + static_cast<size_t>(F) * F, *g_wa = g_bf4 + F;
  float* g_ba = g_wa + static_cast<size_t>(A) * F;
  const uint8_t* keep = t.drop_p > 0.f ? t.keep_mask : nullptr;
  FVLA_REQUIRE(t.drop_p == 0.f || keep != nullptr, "head_train_step: dropout > 0 needs the keep mask");
  const float inv_keep = 1.0f / (1.0f - t.drop_p);
  auto lin = [&](const float* X, int ldx, const float* W, const float* bias, float* Y, int N, int K) {
    if (B > 8) skinny_nt_kernel<16><<<ceil_div(N, 8), 256, 0, s>>>(X, ldx, W, bias, Y, N, B, N, K);
    else skinny_nt_kernel<8><<<ceil_div(N, 8), 256, 0, s>>>(X, ldx, W, bias, Y, N, B, N, K);
  };
  auto dgrad = [&](const float* dY, const float* W, int ldw, float* dX, int ldx, int N, int k0, int Kc) {
    // dX rows are dense here (ldx == Kc), so the split partials are [split][B][Kc] blocks reduced elementwise
    const int splits = N >= 64 * kDgradSplits ? kDgradSplits : 1;
    const int n_per = ceil_div(ceil_div(N, splits), 64) * 64;
    const size_t stride = b * static_cast<size_t>(Kc);
    dim3 grid(ceil_div(Kc, 128), ceil_div(B, 16), splits);
    skinny_nn_kernel<16><<<grid, 128, 0, s>>>(dY, N, W, ldw, splits > 1 ? dpart : dX, ldx, B, N, k0, Kc, n_per, stride);
    if (splits > 1)
      skinny_reduce_kernel<<<ceil_div(B * Kc, 256), 256, 0, s>>>(dpart, stride, splits, dX, B * Kc);
  };
  auto wgrad = [&](const float* dY, const float* X, int ldx, float* dW, float* db, int N, int K) {
    dim3 grid(ceil_div(K, 256), N);
    skinny_tn_kernel<<<grid, 256, 0, s>>>(dY, N, X, ldx, dW, db, B, N, K);
  };
  const int T = 256;
  // ---- forward ----
  ln_fwd_kernel<<<ceil_div(B, 4), 128, 0, s>>>(t.states, S, t.w.ln_s_w, t.w.ln_s_b, shat, rstd_s, sn, B, S);
  lin(sn, S, static_cast<const float*>(t.w.w_state), t.w.b_state, z1, Hd, S);
  silu_cat_kernel<<<ceil_div(B * KC, T), T, 0, s>>>(t.pooled, z1, cat, B, H, Hd);
  lin(cat, KC, static_cast<const float*>(t.w.w_f0), t.w.b_f0, y0, F, KC);
  ln_fwd_kernel<<<ceil_div(B, 4), 128, 0, s>>>(y0, F, t.w.ln_f_w, t.w.ln_f_b, yhat, rstd_f, yn, B, F);
  silu_drop_kernel<<<ceil_div(B * F, T), T, 0, s>>>(yn, keep, inv_keep, x1, B * F);
  lin(x1, F, static_cast<const float*>(t.w.w_f4), t.w.b_f4, y4, F, F);
  silu_drop_kernel<<<ceil_div(B * F, T), T, 0, s>>>(y4, nullptr, 1.f, x2, B * F);
  lin(x2, F, static_cast<const float*>(t.w.w_act), t.w.b_act, a, A, F);
  mse_kernel<<<1, 256, 0, s>>>(a, t.target, da, t.loss, B * A);
  if (t.actions != nullptr)
    FVLA_CUDA_CHECK(cudaMemcpyAsync(t.actions, a, b * A * 4, cudaMemcpyDeviceToDevice, s));
  // ---- backward ----
  wgrad(da, x2, F, g_wa, g_ba, A, F);
  dgrad(da, static_cast<const float*>(t.w.w_act), F, dtmp, F, A, 0, F);                 // dx2
  silu_bwd_kernel<<<ceil_div(B * F, T), T, 0, s>>>(dtmp, F, y4, nullptr, 1.f, y0, B, F);  // dy4 (y0 is free now)
  float* dy4 = y0;
  wgrad(dy4, x1, F, g_wf4, g_bf4, F, F);
  dgrad(dy4, static_cast<const float*>(t.w.w_f4), F, dtmp, F, F, 0, F);                 // dx1 (after dropout)
  silu_bwd_kernel<<<ceil_div(B * F, T), T, 0, s>>>(dtmp, F, yn, keep, inv_keep, x2, B, F);  // dyn (x2 is free now)
  float* dyn = x2;
  colsum2_kernel<<<ceil_div(F, T), T, 0, s>>>(dyn, yhat, g_lnfw, g_lnfb, B, F);
  ln_bwd_kernel<<<ceil_div(B, 4), 128, 0, s>>>(dyn, yhat, rstd_f, t.w.ln_f_w, y4, B, F);  // dy0 (y4 is free now)
  float* dy0 = y4;
  wgrad(dy0, cat, KC, g_wf0, g_bf0, F, KC);
  dgrad(dy0, static_cast<const float*>(t.w.w_f0), KC, dtmp, Hd, F, H, Hd);              // d silu(z1): columns H.. only
  silu_bwd_kernel<<<ceil_div(B * Hd, T), T, 0, s>>>(dtmp, Hd, z1, nullptr, 1.f, dz1, B, Hd);
  wgrad(dz1, sn, S, g_w1, g_b1, Hd, S);
  dgrad(dz1, static_cast<const float*>(t.w.w_state), S, dsn, S, Hd, 0, S);
  colsum2_kernel<<<ceil_div(S, T), T, 0, s>>>(dsn, shat, g_lnsw, g_lnsb, B, S);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fvla
