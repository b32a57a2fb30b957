// Small token-major kernels around the Qwen2 prefill: RMSNorm, token-embedding gather + LLaVA image
// splice, final-norm + pooling, SwiGLU (unfused fallback form), dtype conversion, RoPE tables.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

// Qwen2RMSNorm (transformers/models/qwen2/modeling_qwen2.py: Qwen2RMSNorm.forward):
//   y = weight * (x * rsqrt(mean(x^2) + eps)), statistics in fp32.
// One CTA of 128 threads per row, two passes over the row (the second hits L1/L2); used for H > 2048.  Keeping the
// row in registers between the passes measured slower here (fewer resident CTAs).
constexpr int RMS_THREADS = 128;
template <typename TI, typename TO>
__global__ void __launch_bounds__(RMS_THREADS)
rmsnorm_kernel(const TI* __restrict__ x, const float* __restrict__ w, TO* __restrict__ out, int H,
               float eps) {
  __shared__ float red[32];
  const size_t row = blockIdx.x;
  const TI* xr = x + row * H;
  float ss = 0.f;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    Vec8<TI> v; v.load(xr + i);
#pragma unroll
    for (int c = 0; c < 8; ++c) ss = fmaf(v.v[c], v.v[c], ss);
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / static_cast<float>(H) + eps);
  TO* orow = out + row * H;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    Vec8<TI> v; v.load(xr + i);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + i));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + i) + 1);
    const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
    Vec8<TO> o;
#pragma unroll
    for (int c = 0; c < 8; ++c) o.v[c] = v.v[c] * rstd * wv[c];
    o.store(orow + i);
  }
}

// Warp-per-row variant for H <= 256 * KEEP: no block barrier, KEEP independent 16-byte loads per lane in flight, the
// row read once and kept in registers.  The CTA-per-row kernel above is latency-bound at H = 896 (a CTA's life is
// one load -> two barriers -> one store); four rows per 128-thread CTA with shuffle-only reductions keep ~4x more
// rows in flight per SM.
template <typename TI, typename TO, int KEEP>
__global__ void __launch_bounds__(128)
rmsnorm_warp_kernel(const TI* __restrict__ x, const float* __restrict__ w, TO* __restrict__ out, int rows, int H,
                    float eps) {
  pdl_sync();
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const TI* xr = x + row * H;
  Vec8<TI> keep[KEEP];
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < KEEP; ++k) {
    const int i = (k * 32 + lane) * 8;
    if (i < H) {
      keep[k].load(xr + i);
#pragma unroll
      for (int c = 0; c < 8; ++c) ss = fmaf(keep[k].v[c], keep[k].v[c], ss);
    }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(ss / static_cast<float>(H) + eps);
  TO* orow = out + row * H;
#pragma unroll
  for (int k = 0; k < KEEP; ++k) {
    const int i = (k * 32 + lane) * 8;
    if (i < H) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + i));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + i) + 1);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      Vec8<TO> o;
#pragma unroll
      for (int c = 0; c < 8; ++c) o.v[c] = keep[k].v[c] * rstd * wv[c];
      o.store(orow + i);
    }
  }
}

// FastViTHD LayerNormChannel without its affine part (folded into the qkv GEMM): warp per row, row kept in registers,
// two-pass variance (mean first) in fp32.
template <typename T, int KEEP>
__global__ void __launch_bounds__(128)
layernorm_rows_kernel(const T* __restrict__ x, T* __restrict__ out, int rows, int C, float eps) {
  const int lane = threadIdx.x & 31;
  const long long row = static_cast<long long>(blockIdx.x) * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * C;
  Vec8<T> keep[KEEP];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < KEEP; ++k) {
    const int i = (k * 32 + lane) * 8;
    if (i < C) {
      keep[k].load(xr + i);
#pragma unroll
      for (int c = 0; c < 8; ++c) s += keep[k].v[c];
    }
  }
  const float mean = warp_sum(s) / static_cast<float>(C);
  float vs = 0.f;
#pragma unroll
  for (int k = 0; k < KEEP; ++k) {
    const int i = (k * 32 + lane) * 8;
    if (i < C) {
#pragma unroll
      for (int c = 0; c < 8; ++c) { const float d = keep[k].v[c] - mean; vs = fmaf(d, d, vs); }
    }
  }
  const float rstd = rsqrtf(warp_sum(vs) / static_cast<float>(C) + eps);
  T* orow = out + row * C;
#pragma unroll
  for (int k = 0; k < KEEP; ++k) {
    const int i = (k * 32 + lane) * 8;
    if (i < C) {
      Vec8<T> o;
#pragma unroll
      for (int c = 0; c < 8; ++c) o.v[c] = (keep[k].v[c] - mean) * rstd;
      o.store(orow + i);
    }
  }
}

template <typename TI, typename TO>
void launch_rmsnorm(const void* x, const float* weight, void* out, int rows, int H, float eps, cudaStream_t stream) {
  const TI* xi = static_cast<const TI*>(x);
  TO* oo = static_cast<TO*>(out);
  if (H <= 1024)
    (void)launch_pdl(rmsnorm_warp_kernel<TI, TO, 4>, dim3(ceil_div(rows, 4)), dim3(128), 0, stream, xi, weight, oo, rows, H, eps);
  else if (H <= 2048)
    (void)launch_pdl(rmsnorm_warp_kernel<TI, TO, 8>, dim3(ceil_div(rows, 4)), dim3(128), 0, stream, xi, weight, oo, rows, H, eps);
  else
    rmsnorm_kernel<TI, TO><<<rows, RMS_THREADS, 0, stream>>>(xi, weight, oo, H, eps);
}

// LLaVA prepare_inputs_labels_for_multimodal, materialised from a per-position plan.
template <typename T, typename TO>
__global__ void embed_splice_kernel(const T* __restrict__ table, const T* __restrict__ img,
                                    int n_img, const int* __restrict__ plan, TO* __restrict__ out,
                                    int T_len, int H) {
  pdl_sync();
  const size_t pos = blockIdx.x;  // b*T + t
  const int b = static_cast<int>(pos / T_len);
  const int code = plan[pos];
  TO* dst = out + pos * H;
  const T* src = nullptr;
  if (code >= 0) src = table + static_cast<size_t>(code) * H;
  else if (code <= -2) src = img + (static_cast<size_t>(b) * n_img + (-2 - code)) * H;
  for (int i = threadIdx.x * 8; i < H; i += blockDim.x * 8) {
    Vec8<T> v;
    if (src) v.load(src + i);
    else {
#pragma unroll
      for (int c = 0; c < 8; ++c) v.v[c] = 0.f;
    }
    Vec8<TO> o;
#pragma unroll
    for (int c = 0; c < 8; ++c) o.v[c] = v.v[c];
    o.store(dst + i);
  }
}

// Final RMSNorm fused with FastVLMBackbone._pool_hidden (fastvlm_adapter.py:337-359).
template <typename T>
__global__ void __launch_bounds__(256)
pool_norm_kernel(const T* __restrict__ hidden, const float* __restrict__ w,
                 const int* __restrict__ pool_idx, const int* __restrict__ lens, int mode,
                 float* __restrict__ pooled, int T_len, int H, float eps) {
  pdl_sync();
  __shared__ float red[32];
  const int b = blockIdx.x;
  float* dst = pooled + static_cast<size_t>(b) * H;
  if (mode == 0) {
    int idx = pool_idx[b];
    idx = idx < 0 ? 0 : (idx >= T_len ? T_len - 1 : idx);
    const T* xr = hidden + (static_cast<size_t>(b) * T_len + idx) * H;
    float ss = 0.f;
    for (int i = threadIdx.x; i < H; i += blockDim.x) { const float v = to_f32(xr[i]); ss = fmaf(v, v, ss); }
    ss = block_sum(ss, red);
    const float rstd = rsqrtf(ss / static_cast<float>(H) + eps);
    for (int i = threadIdx.x; i < H; i += blockDim.x) dst[i] = to_f32(xr[i]) * rstd * w[i];
  } else {
    const int n = lens[b];
    for (int i = threadIdx.x; i < H; i += blockDim.x) dst[i] = 0.f;
    for (int t = 0; t < n; ++t) {
      const T* xr = hidden + (static_cast<size_t>(b) * T_len + t) * H;
      float ss = 0.f;
      for (int i = threadIdx.x; i < H; i += blockDim.x) { const float v = to_f32(xr[i]); ss = fmaf(v, v, ss); }
      ss = block_sum(ss, red);
      const float rstd = rsqrtf(ss / static_cast<float>(H) + eps);
      for (int i = threadIdx.x; i < H; i += blockDim.x) dst[i] += to_f32(xr[i]) * rstd * w[i];
    }
    const float denom = fmaxf(static_cast<float>(n), 1e-6f);
    for (int i = threadIdx.x; i < H; i += blockDim.x) dst[i] = dst[i] / denom;
  }
}

template <typename T>
__global__ void swiglu_kernel(const T* __restrict__ gu, T* __restrict__ out, long long total_vec,
                              int I) {
  // gu row: 2*I interleaved (g0,u0,g1,u1,...); each thread: 16 inputs -> 8 outputs
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  Vec8<T> a, b, r;
  a.load(gu + idx * 16);
  b.load(gu + idx * 16 + 8);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    r.v[c] = silu_precise(a.v[2 * c]) * a.v[2 * c + 1];
    r.v[4 + c] = silu_precise(b.v[2 * c]) * b.v[2 * c + 1];
  }
  r.store(out + idx * 8);
}

// Rotate-half RoPE applied in place to the q and k head slices of the fused qkv buffer
// (transformers apply_rotary_pos_emb; position = row index inside the sample).  One thread rotates 8
// (lo, hi) pairs of one head of one token, so attention can stage K/V with plain 16-byte copies.
template <typename T>
__global__ void rope_inplace_kernel(T* __restrict__ qkv, int ld, int n_tok, int heads, int hd,
                                    const float* __restrict__ cs, const float* __restrict__ sn,
                                    long long total) {
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int half = hd >> 1, vpr = half >> 3;
  const int v = static_cast<int>(idx % vpr);
  const int h = static_cast<int>((idx / vpr) % heads);
  const long long row = idx / (static_cast<long long>(vpr) * heads);
  const int pos = static_cast<int>(row % n_tok);
  T* p = qkv + row * ld + h * hd + v * 8;
  Vec8<T> lo, hi;
  lo.load(p);
  hi.load(p + half);
  const float* c = cs + static_cast<size_t>(pos) * half + v * 8;
  const float* s = sn + static_cast<size_t>(pos) * half + v * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float co = __ldg(c + j), si = __ldg(s + j);
    const float a = lo.v[j], b = hi.v[j];
    lo.v[j] = a * co - b * si;
    hi.v[j] = b * co + a * si;
  }
  lo.store(p);
  hi.store(p + half);
}

template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ s, TD* __restrict__ d, long long n) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) {
    float v;
    if constexpr (sizeof(TS) == 1) v = static_cast<float>(s[i]);
    else v = to_f32(s[i]);
    d[i] = from_f32<TD>(v);
  }
}

// HF default rope: inv_freq[i] = theta^(-2i/d); cos/sin of pos*inv_freq in fp32.
__global__ void rope_table_kernel(float* cos_t, float* sin_t, int T_len, int half, float theta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= T_len * half) return;
  const int pos = idx / half, i = idx % half;
  const float inv_freq = 1.0f / powf(theta, static_cast<float>(2 * i) / static_cast<float>(2 * half));
  const float ang = static_cast<float>(pos) * inv_freq;
  cos_t[idx] = cosf(ang);
  sin_t[idx] = sinf(ang);
}

}  // namespace

int layernorm_rows(int dtype, const void* x, void* out, int rows, int C, float eps, cudaStream_t stream) {
  FVLA_REQUIRE(rows > 0 && C > 0 && C % 8 == 0 && C <= 4096, "layernorm_rows: C must be a multiple of 8, <= 4096");
  const int grid = ceil_div(rows, 4);
  if (dtype == DT_F32) {
    if (C <= 1024) layernorm_rows_kernel<float, 4><<<grid, 128, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(out), rows, C, eps);
    else if (C <= 2048) layernorm_rows_kernel<float, 8><<<grid, 128, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(out), rows, C, eps);
    else layernorm_rows_kernel<float, 16><<<grid, 128, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(out), rows, C, eps);
  } else {
    using B16 = __nv_bfloat16;
    if (C <= 1024) layernorm_rows_kernel<B16, 4><<<grid, 128, 0, stream>>>(static_cast<const B16*>(x), static_cast<B16*>(out), rows, C, eps);
    else if (C <= 2048) layernorm_rows_kernel<B16, 8><<<grid, 128, 0, stream>>>(static_cast<const B16*>(x), static_cast<B16*>(out), rows, C, eps);
    else layernorm_rows_kernel<B16, 16><<<grid, 128, 0, stream>>>(static_cast<const B16*>(x), static_cast<B16*>(out), rows, C, eps);
  }
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int rmsnorm(int dtype, const void* x, const float* weight, void* out, int rows, int H, float eps,
            cudaStream_t stream, int x_f32) {
  FVLA_REQUIRE(H % 8 == 0 && rows > 0, "rmsnorm: H%8");
  if (dtype == DT_F32) launch_rmsnorm<float, float>(x, weight, out, rows, H, eps, stream);
  else if (x_f32) launch_rmsnorm<float, __nv_bfloat16>(x, weight, out, rows, H, eps, stream);  // fp32 stream in, bf16 operand out
  else launch_rmsnorm<__nv_bfloat16, __nv_bfloat16>(x, weight, out, rows, H, eps, stream);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int embed_splice(int dtype, const void* table, const void* img_feats, int n_img, const int* plan,
                 void* out, int B, int T, int H, cudaStream_t stream, int out_f32) {
  FVLA_REQUIRE(H % 8 == 0 && B > 0 && T > 0, "embed_splice: H%8");
  if (dtype == DT_F32)
    (void)launch_pdl(embed_splice_kernel<float, float>, dim3(B * T), dim3(128), 0, stream,
        static_cast<const float*>(table), static_cast<const float*>(img_feats), n_img, plan,
        static_cast<float*>(out), T, H);
  else if (out_f32)
    (void)launch_pdl(embed_splice_kernel<__nv_bfloat16, float>, dim3(B * T), dim3(128), 0, stream,
        static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(img_feats),
        n_img, plan, static_cast<float*>(out), T, H);
  else
    (void)launch_pdl(embed_splice_kernel<__nv_bfloat16, __nv_bfloat16>, dim3(B * T), dim3(128), 0, stream,
        static_cast<const __nv_bfloat16*>(table), static_cast<const __nv_bfloat16*>(img_feats),
        n_img, plan, static_cast<__nv_bfloat16*>(out), T, H);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int pool_norm(int dtype, const void* hidden, const float* norm_w, const int* pool_idx,
              const int* lens, int mode, float* pooled, int B, int T, int H, float eps,
              cudaStream_t stream) {
  if (dtype == DT_F32)
    (void)launch_pdl(pool_norm_kernel<float>, dim3(B), dim3(256), 0, stream, static_cast<const float*>(hidden), norm_w,
                                                   pool_idx, lens, mode, pooled, T, H, eps);
  else
    (void)launch_pdl(pool_norm_kernel<__nv_bfloat16>, dim3(B), dim3(256), 0, stream,
        static_cast<const __nv_bfloat16*>(hidden), norm_w, pool_idx, lens, mode, pooled, T, H, eps);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int swiglu_interleaved(int dtype, const void* gu, void* out, int rows, int I, cudaStream_t stream) {
  FVLA_REQUIRE(I % 8 == 0, "swiglu: I%8");
  const long long tv = static_cast<long long>(rows) * (I / 8);
  const unsigned blocks = static_cast<unsigned>(ceil_div_ll(tv, 256));
  if (dtype == DT_F32)
    swiglu_kernel<float><<<blocks, 256, 0, stream>>>(static_cast<const float*>(gu),
                                                     static_cast<float*>(out), tv, I);
  else
    swiglu_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(gu), static_cast<__nv_bfloat16*>(out), tv, I);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int convert(int sd, const void* src, int dd, void* dst, long long n, cudaStream_t s) {
  if (n <= 0) return 0;
  const unsigned blocks = static_cast<unsigned>(ceil_div_ll(n, 256));
#define FVLA_CVT(TS, TD)                                                                      \
  convert_kernel<TS, TD><<<blocks, 256, 0, s>>>(static_cast<const TS*>(src), static_cast<TD*>(dst), n)
  if (sd == DT_F32 && dd == DT_F32) FVLA_CVT(float, float);
  else if (sd == DT_F32 && dd == DT_BF16) FVLA_CVT(float, __nv_bfloat16);
  else if (sd == DT_BF16 && dd == DT_F32) FVLA_CVT(__nv_bfloat16, float);
  else if (sd == DT_BF16 && dd == DT_BF16) FVLA_CVT(__nv_bfloat16, __nv_bfloat16);
  else if (sd == DT_U8 && dd == DT_F32) FVLA_CVT(uint8_t, float);
  else if (sd == DT_U8 && dd == DT_BF16) FVLA_CVT(uint8_t, __nv_bfloat16);
  else { set_error("convert: unsupported dtype pair"); return 2; }
#undef FVLA_CVT
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int rope_inplace(int dtype, void* qkv, int ld, int B, int n_tok, int heads, int head_dim,
                 const float* cos_t, const float* sin_t, cudaStream_t s) {
  FVLA_REQUIRE(head_dim % 16 == 0 && ld % 8 == 0, "rope: head_dim % 16, pitch % 8");
  const long long total = static_cast<long long>(B) * n_tok * heads * (head_dim / 16);
  const unsigned blocks = static_cast<unsigned>(ceil_div_ll(total, 256));
  if (dtype == DT_F32)
    rope_inplace_kernel<float><<<blocks, 256, 0, s>>>(static_cast<float*>(qkv), ld, n_tok, heads, head_dim,
                                                      cos_t, sin_t, total);
  else
    rope_inplace_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(static_cast<__nv_bfloat16*>(qkv), ld, n_tok,
                                                              heads, head_dim, cos_t, sin_t, total);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int rope_table(float* cos_t, float* sin_t, int T, int head_dim, float theta, cudaStream_t s) {
  const int half = head_dim / 2;
  rope_table_kernel<<<ceil_div(T * half, 256), 256, 0, s>>>(cos_t, sin_t, T, half, theta);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fvla
