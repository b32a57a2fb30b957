// fp32 GEMM (FFMA) with the same epilogue contract as the tcgen05 bf16 GEMM.
// Used by the fp32 parity mode (BASELINE config 1: actions within 1e-3 max-abs of the fp32 oracle),
// where single-pass bf16/tf32 tensor-core math does not hold the tolerance through ~110 conv blocks
// and 24 decoder layers.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

constexpr int TM = 128, TN = 64, TK = 16;  // CTA tile; 256 threads, 8x4 outputs per thread

__global__ void __launch_bounds__(256)
gemm_f32_kernel(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                float* D, int ldd, int M, int N, int K,
                const float* __restrict__ bias, const float* __restrict__ row_scale,
                const float* resid, int ldr, int act, int swiglu) {
  __shared__ __align__(16) float As[TK][TM + 4];
  __shared__ __align__(16) float Ws[TK][TN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15;   // 16 column groups of 4
  const int ty = tid >> 4;   // 16 row groups of 8
  const int m0 = blockIdx.x * TM, n0 = blockIdx.y * TN;

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < K; k0 += TK) {
    // A tile: 128 rows x 16 k = 512 float4, two per thread
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int idx = tid + it * 256;
      const int r = idx >> 2, kq = (idx & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int gm = m0 + r, gk = k0 + kq;
      if (gm < M) {
        const float* p = A + static_cast<size_t>(gm) * lda + gk;
        if (gk + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
          if (gk < K) v.x = p[0];
          if (gk + 1 < K) v.y = p[1];
          if (gk + 2 < K) v.z = p[2];
        }
      }
      As[kq][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    }
    {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      const int gn = n0 + r, gk = k0 + kq;
      if (gn < N) {
        const float* p = W + static_cast<size_t>(gn) * ldw + gk;
        if (gk + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
        else {
          if (gk < K) v.x = p[0];
          if (gk + 1 < K) v.y = p[1];
          if (gk + 2 < K) v.z = p[2];
        }
      }
      Ws[kq][r] = v.x; Ws[kq + 1][r] = v.y; Ws[kq + 2][r] = v.z; Ws[kq + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + ty * 8 + i;
    if (m >= M) continue;
    const float rs = row_scale ? row_scale[m] : 1.f;
    const int nb = n0 + tx * 4;
    if (swiglu) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int n = nb + 2 * j;
        if (n + 1 < N) {
          const float g = acc[i][2 * j] * rs, u = acc[i][2 * j + 1] * rs;
          D[static_cast<size_t>(m) * ldd + (n >> 1)] = silu_precise(g) * u;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int n = nb + j;
        if (n < N) {
          float v = acc[i][j] * rs;
          if (bias) v += bias[n];
          v = apply_act_rt(v, act);
          if (resid) v += resid[static_cast<size_t>(m) * ldr + n];
          D[static_cast<size_t>(m) * ldd + n] = v;
        }
      }
    }
  }
}

}  // namespace

int gemm_f32(const GemmArgs& g, cudaStream_t stream) {
  FVLA_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "empty GEMM");
  FVLA_REQUIRE(g.lda % 4 == 0 && g.ldw % 4 == 0, "fp32 GEMM needs 16-byte aligned rows");
  dim3 grid(ceil_div(g.M, TM), ceil_div(g.N, TN));
  gemm_f32_kernel<<<grid, 256, 0, stream>>>(
      static_cast<const float*>(g.A), g.lda, static_cast<const float*>(g.W), g.ldw,
      static_cast<float*>(g.D), g.ldd, g.M, g.N, g.K, g.bias, g.row_scale,
      static_cast<const float*>(g.resid), g.ldr, g.act, g.swiglu);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fvla
