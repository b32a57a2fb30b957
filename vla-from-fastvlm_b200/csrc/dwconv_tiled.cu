// Shared-memory tiled depthwise k x k convolution (stride 1) for the wide FastViTHD stages
// (RepMixer 3x3 and ConvFFN 7x7 at 256^2 x 96, 128^2 x 192, 64^2 x 384): NHWC bf16 in/out, fp32 math.
//
// A CTA (8 warps) produces an 8-row x 64-pixel x 32-channel output tile.  The (8+k-1) x (64+k-1)
// input halo tile is staged once with cp.async (zero-filled outside the image) in a
// [row][channel-vector][x] layout whose x index is padded by one slot every 8 pixels and whose
// channel-vector stride is = 2 (mod 8) 16-byte slots, so every quarter-warp LDS.128 hits 8 distinct
// bank groups.  A warp owns one output row; lane = (channel vector, pixel group): each thread slides
// an 8-pixel window across the row for its 8 channels, keeping one kernel row of weights in
// registers — 49 FMAs per output with ~0.44 shared-memory loads per FMA-group instead of one global
// load per tap.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

constexpr int TW = 64;   // output pixels per tile row
constexpr int TH = 8;    // output rows per tile (= warps per CTA)
constexpr int CB = 32;   // channels per tile (4 vectors of 8)

template <int K> struct TileGeom {
  static constexpr int IW = TW + K - 1;                 // input columns incl. halo
  static constexpr int IH = TH + K - 1;
  static constexpr int XP_RAW = IW + (IW >> 3) + 1;     // padded slots per (row, cvec) line
  static constexpr int XP = XP_RAW + ((2 - (XP_RAW & 7)) & 7);  // = 2 (mod 8)
  static constexpr int IN_BYTES = IH * 4 * XP * 16;
  static constexpr int W_BYTES = K * K * CB * 4;
  static constexpr int SMEM = IN_BYTES + W_BYTES;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes)
               : "memory");
}

template <int K>
__global__ void __launch_bounds__(256, 2)
dwconv_tiled_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                    const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W,
                    int C, int act) {
  using G = TileGeom<K>;
  constexpr int PAD = K / 2;
  extern __shared__ __align__(16) uint8_t smem_dw[];
  const uint32_t s_in = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dw));
  float* s_w = reinterpret_cast<float*>(smem_dw + G::IN_BYTES);

  const int tiles_x = W / TW;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c0 = blockIdx.y * CB;
  const int b = blockIdx.z;
  const int x0 = tx * TW, y0 = ty * TH;
  const int tid = threadIdx.x;

  // ---- stage the halo tile (coalesced 64-byte runs per pixel) ----
  const __nv_bfloat16* img = in + static_cast<size_t>(b) * H * W * C + c0;
  for (int idx = tid; idx < G::IH * G::IW * 4; idx += 256) {
    const int cv = idx & 3;
    const int xi = (idx >> 2) % G::IW;
    const int r = (idx >> 2) / G::IW;
    const int gy = y0 + r - PAD, gx = x0 + xi - PAD;
    const bool ok = gy >= 0 && gy < H && gx >= 0 && gx < W;
    const __nv_bfloat16* src = ok ? img + (static_cast<size_t>(gy) * W + gx) * C + cv * 8 : img;
    const uint32_t dst = s_in + static_cast<uint32_t>(((r * 4 + cv) * G::XP + xi + (xi >> 3)) * 16);
    cp_async16(dst, src, ok ? 16 : 0);
  }
  for (int idx = tid; idx < K * K * CB; idx += 256)
    s_w[idx] = __ldg(w + static_cast<size_t>(idx / CB) * C + c0 + (idx % CB));
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- compute: warp = output row, lane = (channel vector, 8-pixel group) ----
  const int row = tid >> 5, lane = tid & 31;
  const int cv = lane & 3, pg = lane >> 2;
  float acc[8][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + cv * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + cv * 8) + 1);
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    float wr[K][8];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 w0 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * CB + cv * 8);
      const float4 w1 = *reinterpret_cast<const float4*>(s_w + (ky * K + kx) * CB + cv * 8 + 4);
      wr[kx][0] = w0.x; wr[kx][1] = w0.y; wr[kx][2] = w0.z; wr[kx][3] = w0.w;
      wr[kx][4] = w1.x; wr[kx][5] = w1.y; wr[kx][6] = w1.z; wr[kx][7] = w1.w;
    }
    const uint32_t line = s_in + static_cast<uint32_t>((((row + ky) * 4 + cv) * G::XP) * 16);
#pragma unroll
    for (int i = 0; i < 8 + K - 1; ++i) {
      const int xi = pg * 8 + i;
      uint4 raw;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(raw.x), "=r"(raw.y), "=r"(raw.z), "=r"(raw.w)
                   : "r"(line + static_cast<uint32_t>((xi + (xi >> 3)) * 16)));
      float x[8];
      x[0] = __uint_as_float(raw.x << 16); x[1] = __uint_as_float(raw.x & 0xffff0000u);
      x[2] = __uint_as_float(raw.y << 16); x[3] = __uint_as_float(raw.y & 0xffff0000u);
      x[4] = __uint_as_float(raw.z << 16); x[5] = __uint_as_float(raw.z & 0xffff0000u);
      x[6] = __uint_as_float(raw.w << 16); x[7] = __uint_as_float(raw.w & 0xffff0000u);
#pragma unroll
      for (int o = 0; o < 8; ++o) {
        const int kx = i - o;
        if (kx < 0 || kx >= K) continue;
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[o][c] = fmaf(x[c], wr[kx][c], acc[o][c]);
      }
    }
  }
  // ---- store ----
  __nv_bfloat16* orow =
      out + ((static_cast<size_t>(b) * H + (y0 + row)) * W + x0 + pg * 8) * C + c0 + cv * 8;
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    Vec8<__nv_bfloat16> r;
#pragma unroll
    for (int c = 0; c < 8; ++c) r.v[c] = acc[o][c];
    if (act == ACT_GELU) {  // uniform branch: the activation must not be evaluated when unused
#pragma unroll
      for (int c = 0; c < 8; ++c) r.v[c] = gelu_tanh_fit(r.v[c]);
    }
    r.store(orow + static_cast<size_t>(o) * C);
  }
}

template <int K>
int launch_tiled(const void* in, const float* w, const float* bias, void* out, int B, int H, int W,
                 int C, int act, cudaStream_t stream) {
  using G = TileGeom<K>;
  auto kfn = dwconv_tiled_kernel<K>;
  static bool attr_set = false;
  if (!attr_set) {
    FVLA_CUDA_CHECK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, G::SMEM));
    attr_set = true;
  }
  dim3 grid((W / TW) * (H / TH), C / CB, B);
  kfn<<<grid, 256, G::SMEM, stream>>>(static_cast<const __nv_bfloat16*>(in), w, bias,
                                      static_cast<__nv_bfloat16*>(out), H, W, C, act);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

bool dwconv_tiled_supported(int dtype, int H, int W, int C, int mult, int k, int stride) {
  return dtype == DT_BF16 && stride == 1 && mult == 1 && (k == 3 || k == 7) && W % TW == 0 &&
         H % TH == 0 && C % CB == 0;
}

int dwconv_tiled(const void* in, const float* w, const float* bias, void* out, int B, int H, int W,
                 int C, int k, int act, cudaStream_t stream) {
  if (k == 7) return launch_tiled<7>(in, w, bias, out, B, H, W, C, act, stream);
  return launch_tiled<3>(in, w, bias, out, B, H, W, C, act, stream);
}

}  // namespace fvla
