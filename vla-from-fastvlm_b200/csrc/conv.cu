// HBM-bound NHWC kernels of the FastViTHD encoder: image ingest (bilinear letterbox), the dense
// 3-channel stem conv, depthwise / grouped k x k convolutions (RepMixer 3x3, ConvFFN 7x7, RepCPE 7x7,
// patch-embed 7x7 stride 2 with channel multiplier 2, conv_exp 3x3 x2) and the squeeze-excite tail.
// Every thread owns 8 consecutive output channels (one 128-bit bf16 vector) and a short run of
// output pixels, so global traffic is fully vectorised and each input vector is reused across the
// horizontal taps from registers; vertical reuse comes from L1/L2.
#include "common.cuh"
#include "kernels.h"

#include <cstdlib>

namespace fvla {
namespace {

// ------------------------------------------------------------------------------------------
// Image ingest: restates FastVLMBackbone._prepare_images_tensor (fastvlm_adapter.py:479-488):
// channel fix-up (:444-449), aspect-preserving bilinear resize with align_corners=False and
// int() truncated target size, left/top padding (:36-55), optional mean/std (:463-477).
// Output is (B,S,S,4) with a zero 4th channel so the stem reads one aligned vector per pixel.
// ------------------------------------------------------------------------------------------
struct PreGeom {
  int B, C, h, w, S;
  int rh, rw;          // resized extent
  int pad_top, pad_left;
  float sy, sx;        // source / destination scale
  int nhwc;
  float pad_value, scale;
  int normalize;
  float mean[3], inv_std[3];
};

template <typename TS> __device__ __forceinline__ float ld_src(const TS* p) { return to_f32(*p); }
template <> __device__ __forceinline__ float ld_src<uint8_t>(const uint8_t* p) {
  return static_cast<float>(*p);
}

// One thread = PRE_PX consecutive output pixels of one row: the row's vertical interpolation set-up (source rows,
// weights, base pointers) is computed once, and neighbouring outputs re-read the same source texels from L1.
constexpr int PRE_PX = 4;

template <typename TS, typename TD>
__global__ void __launch_bounds__(128)
preprocess_kernel(const TS* __restrict__ src, TD* __restrict__ dst, PreGeom g) {
  pdl_sync();
  // grid (ceil(S / (128 * PRE_PX)), S, B): no index division
  const int xb = (static_cast<int>(blockIdx.x) * blockDim.x + threadIdx.x) * PRE_PX;
  const int y = static_cast<int>(blockIdx.y);
  const int b = static_cast<int>(blockIdx.z);
  if (xb >= g.S) return;
  const int ry = y - g.pad_top;
  const bool row_in = ry >= 0 && ry < g.rh;
  int y0 = 0, y1 = 0;
  float ly = 0.f, hy = 1.f;
  if (row_in) {
    float fy = g.sy * (static_cast<float>(ry) + 0.5f) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    y0 = static_cast<int>(fy);
    y0 = y0 > g.h - 1 ? g.h - 1 : y0;
    y1 = y0 + (y0 < g.h - 1 ? 1 : 0);
    ly = fy - static_cast<float>(y0);
    hy = 1.f - ly;
  }
  const size_t plane = static_cast<size_t>(g.h) * g.w;
  const TS* img = src + static_cast<size_t>(b) * plane * g.C;
  const size_t pix = g.nhwc ? static_cast<size_t>(g.C) : 1;          // element stride between neighbouring texels
  const size_t row0 = static_cast<size_t>(y0) * g.w * pix, row1 = static_cast<size_t>(y1) * g.w * pix;
  float out[PRE_PX][3];
#pragma unroll
  for (int i = 0; i < PRE_PX; ++i) {
    const int x = xb + i;
    const int rx = x - g.pad_left;
    if (!row_in || rx < 0 || rx >= g.rw || x >= g.S) {
      out[i][0] = out[i][1] = out[i][2] = g.pad_value;
    } else {
      float fx = g.sx * (static_cast<float>(rx) + 0.5f) - 0.5f;
      fx = fx < 0.f ? 0.f : fx;
      int x0 = static_cast<int>(fx);
      x0 = x0 > g.w - 1 ? g.w - 1 : x0;
      const int x1 = x0 + (x0 < g.w - 1 ? 1 : 0);
      const float lx = fx - static_cast<float>(x0);
      const float hx = 1.f - lx;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int cs = g.C == 1 ? 0 : c;  // grey -> replicate; >3 channels -> first three
        const TS* base = img + (g.nhwc ? static_cast<size_t>(cs) : static_cast<size_t>(cs) * plane);
        const float v00 = ld_src(base + row0 + static_cast<size_t>(x0) * pix);
        const float v01 = ld_src(base + row0 + static_cast<size_t>(x1) * pix);
        const float v10 = ld_src(base + row1 + static_cast<size_t>(x0) * pix);
        const float v11 = ld_src(base + row1 + static_cast<size_t>(x1) * pix);
        out[i][c] = hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11);
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = out[i][c] * g.scale;
      if (g.normalize) v = (v - g.mean[c]) * g.inv_std[c];
      out[i][c] = v;
    }
  }
  const size_t idx = (static_cast<size_t>(b) * g.S + y) * g.S + xb;
  TD* d = dst + idx * 4;
  if constexpr (sizeof(TD) == 4) {
#pragma unroll
    for (int i = 0; i < PRE_PX; ++i)
      if (xb + i < g.S) reinterpret_cast<float4*>(d)[i] = make_float4(out[i][0], out[i][1], out[i][2], 0.f);
  } else {
    uint32_t w[2 * PRE_PX];
#pragma unroll
    for (int i = 0; i < PRE_PX; ++i) {
      __nv_bfloat162 a = __floats2bfloat162_rn(out[i][0], out[i][1]);
      __nv_bfloat162 c = __floats2bfloat162_rn(out[i][2], 0.f);
      w[2 * i] = *reinterpret_cast<uint32_t*>(&a);
      w[2 * i + 1] = *reinterpret_cast<uint32_t*>(&c);
    }
    if (xb + PRE_PX <= g.S) {  // 32 contiguous bytes
      reinterpret_cast<uint4*>(d)[0] = make_uint4(w[0], w[1], w[2], w[3]);
      reinterpret_cast<uint4*>(d)[1] = make_uint4(w[4], w[5], w[6], w[7]);
    } else {
#pragma unroll
      for (int i = 0; i < PRE_PX; ++i)
        if (xb + i < g.S) reinterpret_cast<uint2*>(d)[i] = make_uint2(w[2 * i], w[2 * i + 1]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// Stem conv: dense 3x3, stride 2, pad 1, 3(+1 zero) input channels -> Cout, + bias + GELU.
// ------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ void ld_px4(const T* p, float (&v)[3]);
template <> __device__ __forceinline__ void ld_px4<float>(const float* p, float (&v)[3]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = a.x; v[1] = a.y; v[2] = a.z;
}
template <> __device__ __forceinline__ void ld_px4<__nv_bfloat16>(const __nv_bfloat16* p,
                                                                   float (&v)[3]) {
  const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
  v[0] = __uint_as_float(a.x << 16);
  v[1] = __uint_as_float(a.x & 0xffff0000u);
  v[2] = __uint_as_float(a.y << 16);
}

template <typename T, int PX>
__global__ void __launch_bounds__(256)
stem_conv_kernel(const T* __restrict__ in, const float* __restrict__ w,
                 const float* __restrict__ bias, T* __restrict__ out, int B, int H, int W,
                 int Cout) {
  const int Ho = H / 2, Wo = W / 2;
  const int CV = Cout / 8, WG = (Wo + PX - 1) / PX;
  // 32-bit index arithmetic (the launcher checks the range): four 64-bit divisions per thread cost more than the
  // 3x3 stencil itself
  const unsigned total = static_cast<unsigned>(B) * Ho * WG * CV;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = static_cast<int>(idx % CV);
  unsigned t = idx / CV;
  const int xg = static_cast<int>(t % WG); t /= WG;
  const int oy = static_cast<int>(t % Ho);
  const int b = static_cast<int>(t / Ho);
  const int ox0 = xg * PX, co0 = cv * 8;

  float acc[PX][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + co0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + co0) + 1);
#pragma unroll
    for (int o = 0; o < PX; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
  constexpr int NCOL = (PX - 1) * 2 + 3;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
    if (iy < 0 || iy >= H) continue;
    const T* row = in + (static_cast<size_t>(b) * H + iy) * W * 4;
#pragma unroll
    for (int ic = 0; ic < NCOL; ++ic) {
      const int ix = ox0 * 2 + ic - 1;
      if (ix < 0 || ix >= W) continue;
      float px[3];
      ld_px4<T>(row + static_cast<size_t>(ix) * 4, px);
#pragma unroll
      for (int o = 0; o < PX; ++o) {
        const int kx = ic - o * 2;
        if (kx < 0 || kx >= 3) continue;
#pragma unroll
        for (int ci = 0; ci < 3; ++ci) {
          const float* wp = w + static_cast<size_t>((ky * 3 + kx) * 3 + ci) * Cout + co0;
          const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
          const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp) + 1);
          const float xv = px[ci];
          acc[o][0] = fmaf(xv, w0.x, acc[o][0]); acc[o][1] = fmaf(xv, w0.y, acc[o][1]);
          acc[o][2] = fmaf(xv, w0.z, acc[o][2]); acc[o][3] = fmaf(xv, w0.w, acc[o][3]);
          acc[o][4] = fmaf(xv, w1.x, acc[o][4]); acc[o][5] = fmaf(xv, w1.y, acc[o][5]);
          acc[o][6] = fmaf(xv, w1.z, acc[o][6]); acc[o][7] = fmaf(xv, w1.w, acc[o][7]);
        }
      }
    }
  }
#pragma unroll
  for (int o = 0; o < PX; ++o) {
    const int ox = ox0 + o;
    if (ox >= Wo) continue;
    Vec8<T> r;
#pragma unroll
    for (int c = 0; c < 8; ++c) r.v[c] = gelu_store<T>(acc[o][c]);
    r.store(out + ((static_cast<size_t>(b) * Ho + oy) * Wo + ox) * Cout + co0);
  }
}

// ------------------------------------------------------------------------------------------
// Depthwise / grouped conv. Thread: 8 output channels (= 8/MULT input channels) x PX pixels.
// ------------------------------------------------------------------------------------------
template <typename T, int MULT> struct InVec;  // loads 8/MULT input channels, expands to 8 lanes
template <typename T> struct InVec<T, 1> {
  __device__ static __forceinline__ void load(const T* p, float (&x)[8]) {
    Vec8<T> v; v.load(p);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = v.v[i];
  }
};
template <> struct InVec<float, 2> {
  __device__ static __forceinline__ void load(const float* p, float (&x)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    x[0] = x[1] = a.x; x[2] = x[3] = a.y; x[4] = x[5] = a.z; x[6] = x[7] = a.w;
  }
};
template <> struct InVec<__nv_bfloat16, 2> {
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&x)[8]) {
    const uint2 a = __ldg(reinterpret_cast<const uint2*>(p));
    x[0] = x[1] = __uint_as_float(a.x << 16);
    x[2] = x[3] = __uint_as_float(a.x & 0xffff0000u);
    x[4] = x[5] = __uint_as_float(a.y << 16);
    x[6] = x[7] = __uint_as_float(a.y & 0xffff0000u);
  }
};

template <typename T, int K, int STRIDE, int MULT, int PX>
__global__ void __launch_bounds__(256)
dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w,
              const float* __restrict__ bias, T* __restrict__ out, int B, int H, int W, int Cin,
              int Ho, int Wo, int act) {
  pdl_sync();
  const int Cout = Cin * MULT;
  const int CV = Cout / 8, WG = (Wo + PX - 1) / PX;
  // 32-bit index arithmetic (the launcher checks the range): four 64-bit divisions per thread cost more than the
  // 3x3 stencil itself
  const unsigned total = static_cast<unsigned>(B) * Ho * WG * CV;
  const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cv = static_cast<int>(idx % CV);
  unsigned t = idx / CV;
  const int xg = static_cast<int>(t % WG); t /= WG;
  const int oy = static_cast<int>(t % Ho);
  const int b = static_cast<int>(t / Ho);
  const int ox0 = xg * PX, co0 = cv * 8, ci0 = co0 / MULT;
  constexpr int PAD = K / 2;

  float acc[PX][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + co0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + co0) + 1);
#pragma unroll
    for (int o = 0; o < PX; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
  constexpr int NCOL = (PX - 1) * STRIDE + K;
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    const int iy = oy * STRIDE + ky - PAD;
    if (iy < 0 || iy >= H) continue;
    float wr[K][8];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float* wp = w + static_cast<size_t>(ky * K + kx) * Cout + co0;
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(wp));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(wp) + 1);
      wr[kx][0] = w0.x; wr[kx][1] = w0.y; wr[kx][2] = w0.z; wr[kx][3] = w0.w;
      wr[kx][4] = w1.x; wr[kx][5] = w1.y; wr[kx][6] = w1.z; wr[kx][7] = w1.w;
    }
    const T* row = in + (static_cast<size_t>(b) * H + iy) * W * Cin + ci0;
#pragma unroll
    for (int ic = 0; ic < NCOL; ++ic) {
      const int ix = ox0 * STRIDE + ic - PAD;
      if (ix < 0 || ix >= W) continue;
      float x[8];
      InVec<T, MULT>::load(row + static_cast<size_t>(ix) * Cin, x);
#pragma unroll
      for (int o = 0; o < PX; ++o) {
        const int kx = ic - o * STRIDE;
        if (kx < 0 || kx >= K) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[o][c] = fmaf(x[c], wr[kx][c], acc[o][c]);
      }
    }
  }
#pragma unroll
  for (int o = 0; o < PX; ++o) {
    const int ox = ox0 + o;
    if (ox >= Wo) continue;
    Vec8<T> r;
#pragma unroll
    for (int c = 0; c < 8; ++c) r.v[c] = acc[o][c];
    if (act == ACT_GELU) {  // uniform branch: the activation must not be evaluated when unused
#pragma unroll
      for (int c = 0; c < 8; ++c) r.v[c] = gelu_store<T>(r.v[c]);
    }
    r.store(out + ((static_cast<size_t>(b) * Ho + oy) * Wo + ox) * Cout + co0);
  }
}

template <typename T, int K, int STRIDE, int MULT, int PX>
int launch_dwconv(const void* in, const float* w, const float* bias, void* out, int B, int H, int W,
                  int Cin, int act, cudaStream_t stream) {
  const int Ho = (H + 2 * (K / 2) - K) / STRIDE + 1, Wo = (W + 2 * (K / 2) - K) / STRIDE + 1;
  const int Cout = Cin * MULT;
  const long long total =
      static_cast<long long>(B) * Ho * ((Wo + PX - 1) / PX) * (Cout / 8);
  const int threads = 256;
  const long long blocks = ceil_div_ll(total, threads);
  FVLA_REQUIRE(total < (1ll << 32) - threads, "dwconv: too many work items for 32-bit indexing");
  (void)launch_pdl(dwconv_kernel<T, K, STRIDE, MULT, PX>, dim3(static_cast<unsigned>(blocks)), dim3(threads), 0, stream,
      static_cast<const T*>(in), w, bias, static_cast<T*>(out), B, H, W, Cin, Ho, Wo, act);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <typename T>
int dwconv_dispatch(const void* in, const float* w, const float* bias, void* out, int B, int H,
                    int W, int Cin, int mult, int k, int stride, int act, cudaStream_t s) {
  if (k == 3 && stride == 1 && mult == 1) return launch_dwconv<T, 3, 1, 1, 4>(in, w, bias, out, B, H, W, Cin, act, s);
  if (k == 3 && stride == 2 && mult == 1) return launch_dwconv<T, 3, 2, 1, 2>(in, w, bias, out, B, H, W, Cin, act, s);
  if (k == 3 && stride == 1 && mult == 2) return launch_dwconv<T, 3, 1, 2, 4>(in, w, bias, out, B, H, W, Cin, act, s);
  if (k == 7 && stride == 1 && mult == 1) return launch_dwconv<T, 7, 1, 1, 4>(in, w, bias, out, B, H, W, Cin, act, s);
  if (k == 7 && stride == 2 && mult == 2) return launch_dwconv<T, 7, 2, 2, 4>(in, w, bias, out, B, H, W, Cin, act, s);
  if (k == 7 && stride == 2 && mult == 1) return launch_dwconv<T, 7, 2, 1, 4>(in, w, bias, out, B, H, W, Cin, act, s);
  set_error("dwconv: unsupported (k, stride, mult) = (" + std::to_string(k) + ", " +
            std::to_string(stride) + ", " + std::to_string(mult) + ")");
  return 2;
}

// ------------------------------------------------------------------------------------------
// Squeeze-excite + GELU (conv_exp tail). mean over HW -> 2 tiny FCs -> sigmoid -> scale -> GELU
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
se_mean_kernel(const T* __restrict__ x, float* __restrict__ mean, int HW, int C) {
  pdl_sync();
  // grid (ceil(C/64), B); 256 threads = 8 channel vectors x 32 pixel lanes; pixels strided by 32
  __shared__ float part[32][65];
  const int cvl = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int c0 = blockIdx.x * 64 + cvl * 8;
  const int b = blockIdx.y;
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 < C) {
    const T* p = x + static_cast<size_t>(b) * HW * C + c0;
    for (int i = pl; i < HW; i += 32) {
      Vec8<T> v; v.load(p + static_cast<size_t>(i) * C);
#pragma unroll
      for (int c = 0; c < 8; ++c) s[c] += v.v[c];
    }
  }
#pragma unroll
  for (int c = 0; c < 8; ++c) part[pl][cvl * 8 + c] = s[c];
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) t += part[i][threadIdx.x];
    const int c = blockIdx.x * 64 + threadIdx.x;
    if (c < C) mean[static_cast<size_t>(b) * C + c] = t / static_cast<float>(HW);
  }
}

// hidden[b][r] = relu(W1[r] . mean[b] + b1[r]).  The first version ran one block per sample and streamed both weight
// matrices (4.7 MB) through every block with a stride-Cr access on W2: 0.53 ms at any batch size, 8 % of the b=1
// forward; a version that cached the weight row in 128 registers per lane was occupancy-starved (0.29 ms).
__global__ void __launch_bounds__(256)
se_fc1_kernel(const float* __restrict__ mean, const float* __restrict__ w1, const float* __restrict__ b1,
              float* __restrict__ hidden, int B, int C, int Cr) {
  pdl_sync();
  // one warp per (reduced channel r, sample b): a 3072-long dot product as 128-bit loads, 4 in flight per operand
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int r = blockIdx.x;
  const int b = blockIdx.y * nw + warp;
  if (b >= B) return;
  const float4* wr = reinterpret_cast<const float4*>(w1 + static_cast<size_t>(r) * C);
  const float4* m = reinterpret_cast<const float4*>(mean + static_cast<size_t>(b) * C);
  const int n4 = C >> 2;  // C % 4 == 0 (checked by the launcher)
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = lane;
  for (; i + 96 < n4; i += 128) {
    const float4 w0 = __ldg(wr + i), w1v = __ldg(wr + i + 32), w2 = __ldg(wr + i + 64), w3 = __ldg(wr + i + 96);
    const float4 x0 = m[i], x1 = m[i + 32], x2 = m[i + 64], x3 = m[i + 96];
    a0 = fmaf(w0.x, x0.x, fmaf(w0.y, x0.y, fmaf(w0.z, x0.z, fmaf(w0.w, x0.w, a0))));
    a1 = fmaf(w1v.x, x1.x, fmaf(w1v.y, x1.y, fmaf(w1v.z, x1.z, fmaf(w1v.w, x1.w, a1))));
    a2 = fmaf(w2.x, x2.x, fmaf(w2.y, x2.y, fmaf(w2.z, x2.z, fmaf(w2.w, x2.w, a2))));
    a3 = fmaf(w3.x, x3.x, fmaf(w3.y, x3.y, fmaf(w3.z, x3.z, fmaf(w3.w, x3.w, a3))));
  }
  for (; i < n4; i += 32) {
    const float4 w0 = __ldg(wr + i);
    const float4 x0 = m[i];
    a0 = fmaf(w0.x, x0.x, fmaf(w0.y, x0.y, fmaf(w0.z, x0.z, fmaf(w0.w, x0.w, a0))));
  }
  const float a = warp_sum((a0 + a1) + (a2 + a3));
  if (lane == 0) hidden[static_cast<size_t>(b) * Cr + r] = fmaxf(a + b1[r], 0.f);
}
// gate[b][c] = sigmoid(W2[c] . hidden[b] + b2[c]): one warp per output channel, weight row in registers
__global__ void __launch_bounds__(256)
se_fc2_kernel(const float* __restrict__ hidden, const float* __restrict__ w2, const float* __restrict__ b2,
              float* __restrict__ gate, int B, int C, int Cr) {
  pdl_sync();
  constexpr int MAXV = 8;  // Cr <= 256
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * (blockDim.x >> 5) + warp;
  if (c >= C) return;
  const float* wr = w2 + static_cast<size_t>(c) * Cr;
  float wreg[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i) wreg[i] = (lane + 32 * i < Cr) ? __ldg(wr + lane + 32 * i) : 0.f;
  const float bias = b2[c];
  for (int b = 0; b < B; ++b) {
    const float* h = hidden + static_cast<size_t>(b) * Cr;
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
      if (lane + 32 * i < Cr) a = fmaf(wreg[i], h[lane + 32 * i], a);
    a = warp_sum(a);
    if (lane == 0) gate[static_cast<size_t>(b) * C + c] = 1.f / (1.f + expf(-(a + bias)));
  }
}

template <typename T>
__global__ void se_scale_gelu_kernel(const T* __restrict__ x, const float* __restrict__ gate,
                                     T* __restrict__ out, int HW, int C, long long total_vec) {
  pdl_sync();
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total_vec) return;
  const int CV = C / 8;
  const int cv = static_cast<int>(idx % CV);
  const int b = static_cast<int>(idx / (static_cast<long long>(CV) * HW));
  Vec8<T> v; v.load(x + idx * 8);
  const float* g = gate + static_cast<size_t>(b) * C + cv * 8;
#pragma unroll
  for (int c = 0; c < 8; ++c) v.v[c] = gelu_erf(v.v[c] * g[c]);
  v.store(out + idx * 8);
}

template <typename T>
int se_gelu_t(const void* x, void* out, int B, int HW, int C, int Cr, const float* w1,
              const float* b1, const float* w2, const float* b2, float* mean, float* gate,
              cudaStream_t s) {
  dim3 g1(ceil_div(C, 64), B);
  (void)launch_pdl(se_mean_kernel<T>, g1, dim3(256), 0, s, static_cast<const T*>(x), mean, HW, C);
  // no extra scratch: fc1 writes the hidden vectors [B, Cr] into the gate buffer, fc2 reads them from there and
  // writes the gates [B, C] over the (now consumed) means, which the scale kernel then reads
  FVLA_REQUIRE(C % 4 == 0 && Cr <= 256 && Cr <= C, "se_gelu: channel counts out of range");
  float* hidden = gate;
  float* gates = mean;
  (void)launch_pdl(se_fc1_kernel, dim3(Cr, ceil_div(B, 8)), dim3(256), 0, s, mean, w1, b1, hidden, B, C, Cr);
  (void)launch_pdl(se_fc2_kernel, dim3(ceil_div(C, 8)), dim3(256), 0, s, hidden, w2, b2, gates, B, C, Cr);
  const long long tv = static_cast<long long>(B) * HW * (C / 8);
  (void)launch_pdl(se_scale_gelu_kernel<T>, dim3(static_cast<unsigned>(ceil_div_ll(tv, 256))), dim3(256), 0, s,
      static_cast<const T*>(x), gates, static_cast<T*>(out), HW, C, tv);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <typename TS, typename TD>
int launch_pre(const PreprocessArgs& a, const PreGeom& g, cudaStream_t s) {
  FVLA_REQUIRE(a.S <= 65535 && a.B <= 65535, "preprocess: image side / batch exceed the launch grid");
  dim3 grid(static_cast<unsigned>(ceil_div(a.S, 128 * PRE_PX)), static_cast<unsigned>(a.S), static_cast<unsigned>(a.B));
  (void)launch_pdl(preprocess_kernel<TS, TD>, grid, dim3(128), 0, s, static_cast<const TS*>(a.src), static_cast<TD*>(a.dst), g);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

int preprocess_images(int dtype, const PreprocessArgs& a, cudaStream_t stream) {
  FVLA_REQUIRE(a.B > 0 && a.h > 0 && a.w > 0 && a.S > 0 && a.C >= 1, "bad image geometry");
  PreGeom g;
  g.B = a.B; g.C = a.C; g.h = a.h; g.w = a.w; g.S = a.S; g.nhwc = a.src_nhwc;
  g.pad_value = a.pad_value; g.scale = a.scale; g.normalize = a.normalize;
  for (int i = 0; i < 3; ++i) { g.mean[i] = a.mean[i]; g.inv_std[i] = a.inv_std[i]; }
  if (a.letterbox) {
    // resize_with_pad (fastvlm_adapter.py:44-54): python float division, int() truncation
    const double ratio = std::max(static_cast<double>(a.w) / a.S, static_cast<double>(a.h) / a.S);
    g.rh = static_cast<int>(static_cast<double>(a.h) / ratio);
    g.rw = static_cast<int>(static_cast<double>(a.w) / ratio);
    g.pad_top = std::max(0, a.S - g.rh);
    g.pad_left = std::max(0, a.S - g.rw);
  } else {
    g.rh = a.S; g.rw = a.S; g.pad_top = 0; g.pad_left = 0;
  }
  FVLA_REQUIRE(g.rh >= 1 && g.rw >= 1 && g.rh <= a.S && g.rw <= a.S, "degenerate resize target");
  // torch upsample (size given, align_corners=False): scale = in / out in float
  g.sy = static_cast<float>(a.h) / static_cast<float>(g.rh);
  g.sx = static_cast<float>(a.w) / static_cast<float>(g.rw);
  if (dtype == DT_F32) {
    if (a.src_dtype == DT_F32) return launch_pre<float, float>(a, g, stream);
    if (a.src_dtype == DT_U8) return launch_pre<uint8_t, float>(a, g, stream);
    if (a.src_dtype == DT_BF16) return launch_pre<__nv_bfloat16, float>(a, g, stream);
  } else {
    if (a.src_dtype == DT_F32) return launch_pre<float, __nv_bfloat16>(a, g, stream);
    if (a.src_dtype == DT_U8) return launch_pre<uint8_t, __nv_bfloat16>(a, g, stream);
    if (a.src_dtype == DT_BF16) return launch_pre<__nv_bfloat16, __nv_bfloat16>(a, g, stream);
  }
  set_error("preprocess: unsupported source dtype");
  return 2;
}

int stem_conv3x3_s2(int dtype, const void* in, const float* w_packed, const float* bias, void* out,
                    int B, int H, int W, int Cout, cudaStream_t stream) {
  FVLA_REQUIRE(Cout % 8 == 0 && H % 2 == 0 && W % 2 == 0, "stem conv: Cout%8, even H/W");
  constexpr int PX = 4;
  const long long total = static_cast<long long>(B) * (H / 2) * ceil_div(W / 2, PX) * (Cout / 8);
  const unsigned blocks = static_cast<unsigned>(ceil_div_ll(total, 256));
  if (dtype == DT_F32)
    stem_conv_kernel<float, PX><<<blocks, 256, 0, stream>>>(
        static_cast<const float*>(in), w_packed, bias, static_cast<float*>(out), B, H, W, Cout);
  else
    stem_conv_kernel<__nv_bfloat16, PX><<<blocks, 256, 0, stream>>>(
        static_cast<const __nv_bfloat16*>(in), w_packed, bias, static_cast<__nv_bfloat16*>(out), B,
        H, W, Cout);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int dwconv(int dtype, const void* in, const float* w_packed, const float* bias, void* out, int B,
           int H, int W, int Cin, int mult, int ksize, int stride, int act, cudaStream_t stream,
           const uint32_t* wtab) {
  FVLA_REQUIRE((Cin * mult) % 8 == 0, "dwconv: output channels must be a multiple of 8");
  static const bool mma_on = std::getenv("FVLA_DISABLE_DWCONV_MMA") == nullptr;  // A/B switch (precision / profiling)
  if (mma_on && dwconv7_mma_supported(dtype, H, W, Cin, mult, ksize, stride, act)) {
    if (wtab == nullptr) {
      // op-level callers (tests, micro-benchmarks) pass only the fp32 taps: build the table into a scratch
      // buffer on the same stream (the engine builds its tables once at finalize)
      static uint32_t* scratch = nullptr;
      static size_t scratch_bytes = 0;
      const size_t need = dwconv7_wtab_bytes(Cin);
      if (need > scratch_bytes) {
        FVLA_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (scratch != nullptr) FVLA_CUDA_CHECK(cudaFree(scratch));
        FVLA_CUDA_CHECK(cudaMalloc(&scratch, need));
        scratch_bytes = need;
      }
      if (int rc = dwconv7_mma_prepare(w_packed, Cin, scratch, stream)) return rc;
      wtab = scratch;
    }
    return dwconv7_mma(in, wtab, bias, out, B, H, W, Cin, stream);
  }
  if (mma_on && dwconv7_s2m2_mma_supported(dtype, H, W, Cin, mult, ksize, stride, act)) {
    if (wtab == nullptr) {   // op-level callers: build the table into a scratch buffer on the same stream
      static uint32_t* scratch2 = nullptr;
      static size_t scratch2_bytes = 0;
      const size_t need = dwconv7_wtab_bytes(Cin);
      if (need > scratch2_bytes) {
        FVLA_CUDA_CHECK(cudaStreamSynchronize(stream));
        if (scratch2 != nullptr) FVLA_CUDA_CHECK(cudaFree(scratch2));
        FVLA_CUDA_CHECK(cudaMalloc(&scratch2, need));
        scratch2_bytes = need;
      }
      if (int rc = dwconv7_s2m2_mma_prepare(w_packed, Cin, scratch2, stream)) return rc;
      wtab = scratch2;
    }
    return dwconv7_s2m2_mma(in, wtab, bias, out, B, H, W, Cin, act, stream);
  }
  const bool tma3_on = std::getenv("FVLA_DISABLE_DWCONV3_TMA") == nullptr;  // A/B switch, read per call (tests toggle it)
  if (tma3_on && dwconv3_tma_supported(dtype, H, W, Cin, mult, ksize, stride, act))
    return dwconv3_tma(in, w_packed, bias, out, B, H, W, Cin, stream);
  if (dwconv_s2m2_tiled_supported(dtype, H, W, Cin, mult, ksize, stride) && (act == ACT_NONE || act == ACT_GELU))
    return dwconv_s2m2_tiled(in, w_packed, bias, out, B, H, W, Cin, act, stream);
  static const bool tiled_on = std::getenv("FVLA_DISABLE_DWCONV_TILED") == nullptr;  // A/B switch (precision / profiling)
  if (tiled_on && dwconv_tiled_supported(dtype, H, W, Cin, mult, ksize, stride))
    return dwconv_tiled(in, w_packed, bias, out, B, H, W, Cin, ksize, act, stream);
  if (dtype == DT_F32)
    return dwconv_dispatch<float>(in, w_packed, bias, out, B, H, W, Cin, mult, ksize, stride, act, stream);
  return dwconv_dispatch<__nv_bfloat16>(in, w_packed, bias, out, B, H, W, Cin, mult, ksize, stride, act, stream);
}

int se_gelu(int dtype, const void* x, void* out, int B, int HW, int C, int Cr, const float* w1,
            const float* b1, const float* w2, const float* b2, float* scratch_mean,
            float* scratch_gate, cudaStream_t stream) {
  FVLA_REQUIRE(C % 8 == 0, "se: C%8");
  if (dtype == DT_F32)
    return se_gelu_t<float>(x, out, B, HW, C, Cr, w1, b1, w2, b2, scratch_mean, scratch_gate, stream);
  return se_gelu_t<__nv_bfloat16>(x, out, B, HW, C, Cr, w1, b1, w2, b2, scratch_mean, scratch_gate, stream);
}

}  // namespace fvla
