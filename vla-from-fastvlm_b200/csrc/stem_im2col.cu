// FastViTHD stem.0 (dense 3x3 stride-2 conv, 3 -> d0 channels, + GELU) on the tensor cores:
// the 27-tap patches of the ingested image are gathered into a [B*(S/2)^2, 32] bf16 matrix (27 taps
// + 5 zero columns, one 64-byte row per output pixel) and the conv becomes a K=32 GEMM whose
// bias + GELU run in the tcgen05 epilogue.  Used by the bf16 mode; the fp32 parity mode keeps the
// direct FFMA kernel in conv.cu.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

// in: [B,H,W,4] bf16 (4th channel zero); col: [B*(H/2)*(W/2), 32] with column (ky*3+kx)*3 + ci
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ col, int B,
                   int H, int W) {
  const int Ho = H / 2, Wo = W / 2;
  const long long total = static_cast<long long>(B) * Ho * Wo;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = static_cast<int>(idx % Wo);
  const int oy = static_cast<int>((idx / Wo) % Ho);
  const int b = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
  uint16_t v[32];
#pragma unroll
  for (int i = 27; i < 32; ++i) v[i] = 0;
#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * 2 + ky - 1;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * 2 + kx - 1;
      uint2 px = make_uint2(0u, 0u);
      if (iy >= 0 && iy < H && ix >= 0 && ix < W)
        px = __ldg(reinterpret_cast<const uint2*>(in + ((static_cast<size_t>(b) * H + iy) * W + ix) * 4));
      const int p = (ky * 3 + kx) * 3;
      v[p] = static_cast<uint16_t>(px.x & 0xffffu);
      v[p + 1] = static_cast<uint16_t>(px.x >> 16);
      v[p + 2] = static_cast<uint16_t>(px.y & 0xffffu);
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(col + idx * 32);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 o;
    o.x = v[8 * q] | (static_cast<uint32_t>(v[8 * q + 1]) << 16);
    o.y = v[8 * q + 2] | (static_cast<uint32_t>(v[8 * q + 3]) << 16);
    o.z = v[8 * q + 4] | (static_cast<uint32_t>(v[8 * q + 5]) << 16);
    o.w = v[8 * q + 6] | (static_cast<uint32_t>(v[8 * q + 7]) << 16);
    dst[q] = o;
  }
}

}  // namespace

int stem_im2col_bf16(const void* in, void* col, int B, int H, int W, cudaStream_t stream) {
  FVLA_REQUIRE(H % 2 == 0 && W % 2 == 0, "stem im2col: even H/W");
  const long long total = static_cast<long long>(B) * (H / 2) * (W / 2);
  stem_im2col_kernel<<<static_cast<unsigned>(ceil_div_ll(total, 256)), 256, 0, stream>>>(
      static_cast<const __nv_bfloat16*>(in), static_cast<__nv_bfloat16*>(col), B, H, W);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace fvla
