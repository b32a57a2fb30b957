// Shared-memory tiled depthwise k x k convolution (stride 1) for the wide FastViTHD stages
// (RepMixer 3x3 and ConvFFN 7x7 at 256^2 x 96, 128^2 x 192, 64^2 x 384): NHWC bf16 in/out, fp16 taps
// flushed into fp32 accumulators.
//
// A CTA (8 warps) produces an 8-row x 64-pixel x 32-channel output tile.  The (8+k-1) x (64+k-1)
// input halo tile is staged once (zero-filled outside the image, converted to fp16) in a
// [row][channel-vector][x] layout whose x index is padded by one slot every 8 pixels and whose
// channel-vector stride is = 2 (mod 8) 16-byte slots, so every quarter-warp LDS.128 hits 8 distinct
// bank groups.  A warp owns one output row; lane = (channel vector, pixel group): each thread slides
// an 8-pixel window across the row for its 8 channels, keeping one kernel row of weights in
// registers.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

constexpr int TH = 8;    // output rows per tile (= warps per CTA)
constexpr int CB = 32;   // channels per tile (4 vectors of 8)
// output pixels per tile row: 64 (8 per lane) for 7x7; 32 (4 per lane) for 3x3, whose light arithmetic wants
// occupancy instead of register blocking — half the accumulators and staging registers give 3 CTAs per SM
template <int K> constexpr int tile_w() { return K == 3 ? 32 : 64; }

template <int K> struct TileGeom {
  static constexpr int TW = tile_w<K>();
  static constexpr int PXT = TW / 8;                    // output pixels per lane
  static constexpr int IW = TW + K - 1;                 // input columns incl. halo
  static constexpr int IH = TH + K - 1;
  static constexpr int PSH = PXT == 8 ? 3 : 2;          // one padding slot per lane group: xi + (xi >> PSH)
  static constexpr int XP_RAW = IW + (IW >> PSH) + 1;   // padded slots per (row, cvec) line
  static constexpr int XP = XP_RAW + ((2 - (XP_RAW & 7)) & 7);  // = 2 (mod 8)
  static constexpr int IN_BYTES = IH * 4 * XP * 16;
};

// ---------------------------------------------------------------------------------------------
// Packed-half arithmetic.  A first version of this kernel unpacked bf16 to fp32 per tap and was
// instruction-bound (ncu, profiles/r01_dwconv_k7_c192_ncu.txt: issue slots 64 % busy, FMA pipe 37 %,
// ALU pipe 43 % — unpacking and shared-memory loads cost as many slots as the FMAs; 23 TFLOP/s).
// Here the halo tile is converted ONCE to fp16 while it is staged (bf16 values are
// exactly representable in fp16 inside its range; out-of-range values saturate), the weights are
// kept as fp16, and the taps of two kernel rows are accumulated with HFMA2 (two channels per
// instruction, no unpacking) before being flushed into the fp32 accumulators.  fp16 carries 3 more
// mantissa bits than the bf16 output, so the <= 14-term partial sums add ~2^-11 relative error,
// a quarter of the output rounding.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bf16x2_to_f16x2_sat(uint32_t w) {
  const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t hmul2_u(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float2 h2_to_f2(uint32_t h) {
  float2 f;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n"
      : "=f"(f.x), "=f"(f.y)
      : "r"(h));
  return f;
}

template <int K>
__global__ void __launch_bounds__(256, (K == 3 ? 3 : 2))
dwconv_tiled_h_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W,
                      int C, int act) {
  using G = TileGeom<K>;
  constexpr int TW = G::TW, PXT = G::PXT;
  constexpr int PAD = K / 2;
  extern __shared__ __align__(16) uint8_t smem_dw[];
  const uint32_t s_in = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dw));
  uint32_t* s_wh = reinterpret_cast<uint32_t*>(smem_dw + G::IN_BYTES);  // [K*K][CB/2] half2

  const int tiles_x = W / TW;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int c0 = blockIdx.y * CB;
  const int b = blockIdx.z;
  const int x0 = tx * TW, y0 = ty * TH;
  const int tid = threadIdx.x;

  // ---- stage the halo tile, converting bf16 -> fp16 (coalesced 64-byte runs per pixel) ----
  const __nv_bfloat16* img = in + static_cast<size_t>(b) * H * W * C + c0;
  constexpr int NV = G::IH * G::IW * 4;
  constexpr int PER = (NV + 255) / 256;
  uint4 stage[PER];
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int idx = tid + it * 256;
    stage[it] = make_uint4(0u, 0u, 0u, 0u);
    if (idx < NV) {
      const int cv = idx & 3;
      const int xi = (idx >> 2) % G::IW;
      const int r = (idx >> 2) / G::IW;
      const int gy = y0 + r - PAD, gx = x0 + xi - PAD;
      if (gy >= 0 && gy < H && gx >= 0 && gx < W)
        stage[it] = __ldg(reinterpret_cast<const uint4*>(img + (static_cast<size_t>(gy) * W + gx) * C + cv * 8));
    }
  }
  for (int idx = tid; idx < K * K * (CB / 2); idx += 256) {
    const int t = idx / (CB / 2), cp = idx % (CB / 2);
    const float w0 = __ldg(w + static_cast<size_t>(t) * C + c0 + 2 * cp);
    const float w1 = __ldg(w + static_cast<size_t>(t) * C + c0 + 2 * cp + 1);
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(w1), "f"(w0));
    s_wh[idx] = r;
  }
#pragma unroll
  for (int it = 0; it < PER; ++it) {
    const int idx = tid + it * 256;
    if (idx < NV) {
      const int cv = idx & 3;
      const int xi = (idx >> 2) % G::IW;
      const int r = (idx >> 2) / G::IW;
      const uint32_t dst = s_in + static_cast<uint32_t>(((r * 4 + cv) * G::XP + xi + (xi >> G::PSH)) * 16);
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(bf16x2_to_f16x2_sat(stage[it].x)),
                   "r"(bf16x2_to_f16x2_sat(stage[it].y)), "r"(bf16x2_to_f16x2_sat(stage[it].z)),
                   "r"(bf16x2_to_f16x2_sat(stage[it].w))
                   : "memory");
    }
  }
  __syncthreads();

  // ---- compute: warp = output row, lane = (channel vector, 8-pixel group) ----
  const int row = tid >> 5, lane = tid & 31;
  const int cv = lane & 3, pg = lane >> 2;
  float acc[PXT][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0 + cv * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + cv * 8) + 1);
#pragma unroll
    for (int o = 0; o < PXT; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
  uint32_t racc[PXT][4];  // packed fp16 partial sums of up to two kernel rows
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    uint32_t wr[K][4];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const uint4 wv = *reinterpret_cast<const uint4*>(s_wh + (ky * K + kx) * (CB / 2) + cv * 4);
      wr[kx][0] = wv.x; wr[kx][1] = wv.y; wr[kx][2] = wv.z; wr[kx][3] = wv.w;
    }
    const bool fresh = (ky & 1) == 0;  // first row of a pair: start the partial sums with a multiply
    const uint32_t line = s_in + static_cast<uint32_t>((((row + ky) * 4 + cv) * G::XP) * 16);
#pragma unroll
    for (int i = 0; i < PXT + K - 1; ++i) {
      const int xi = pg * PXT + i;
      uint32_t x[4];
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                   : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3])
                   : "r"(line + static_cast<uint32_t>((xi + (xi >> G::PSH)) * 16)));
#pragma unroll
      for (int o = 0; o < PXT; ++o) {
        const int kx = i - o;
        if (kx < 0 || kx >= K) continue;
        if (kx == 0 && fresh) {
#pragma unroll
          for (int c = 0; c < 4; ++c) racc[o][c] = hmul2_u(x[c], wr[0][c]);
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) racc[o][c] = hfma2_u(x[c], wr[kx][c], racc[o][c]);
        }
      }
    }
    if (!fresh || ky == K - 1) {  // flush the pair into the fp32 accumulators
#pragma unroll
      for (int o = 0; o < PXT; ++o)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float2 f = h2_to_f2(racc[o][c]);
          acc[o][2 * c] += f.x;
          acc[o][2 * c + 1] += f.y;
        }
    }
  }
  // ---- store ----
  __nv_bfloat16* orow =
      out + ((static_cast<size_t>(b) * H + (y0 + row)) * W + x0 + pg * PXT) * C + c0 + cv * 8;
#pragma unroll
  for (int o = 0; o < PXT; ++o) {
    Vec8<__nv_bfloat16> r;
#pragma unroll
    for (int c = 0; c < 8; ++c) r.v[c] = acc[o][c];
    if (act == ACT_GELU) {
#pragma unroll
      for (int c = 0; c < 8; ++c) r.v[c] = gelu_tanh_fit(r.v[c]);
    }
    r.store(orow + static_cast<size_t>(o) * C);
  }
}

template <int K>
int launch_tiled_h(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int C,
                   int act, cudaStream_t stream) {
  using G = TileGeom<K>;
  auto kfn = dwconv_tiled_h_kernel<K>;
  constexpr int SMEM = G::IN_BYTES + K * K * (CB / 2) * 4;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), SMEM)) return rc;
  dim3 grid((W / G::TW) * (H / TH), C / CB, B);
  kfn<<<grid, 256, SMEM, stream>>>(static_cast<const __nv_bfloat16*>(in), w, bias,
                                   static_cast<__nv_bfloat16*>(out), H, W, C, act);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Stride-2, x2-channel grouped 7x7 (FastViTHD patch-embed `proj.0`: groups = Cin, Cout = 2 Cin) + bias + GELU.
// Same scheme as above — halo tile staged once as fp16, packed-half taps flushed into fp32 every two kernel
// rows — with two twists: output pixel o reads inputs 2o + kx (a lane's 4 outputs slide over 13 inputs per
// kernel row), and output channels (2c, 2c+1) both read input channel c, so each loaded value is duplicated
// into both halves of a register and ONE HFMA2 serves the channel pair.
// CTA = 8 output rows x 32 output pixels x 32 output channels; tile [21 rows][4 vecs of 4 in-ch][69 px] fp16.
// ---------------------------------------------------------------------------------------------
// Two tile shapes: 8 rows x 32 px (warp = one output row) for the wide maps, 16 rows x 16 px (warp = two output rows)
// for the 16 x 16 output of the last patch-embed, which used to fall back to the generic kernel (0.35 ms for 0.04 GF).
constexpr int S2_CBO = 32;
template <int TWO> struct S2Geo {
  static constexpr int THO = 256 / TWO;                   // 8 warps x (32 / (TWO / 4)) rows
  static constexpr int IW = 2 * TWO + 5, IH = 2 * THO + 5;
  static constexpr int XP_RAW = IW + (IW >> 3) + 1;
  static constexpr int XP = XP_RAW + ((4 - (XP_RAW & 15)) & 15);  // = 4 (mod 16) 8-byte slots: conflict-free LDS.64
  static constexpr int IN_BYTES = IH * 4 * XP * 8;
  static constexpr int SMEM = IN_BYTES + 49 * (S2_CBO / 2) * 4;
};

__device__ __forceinline__ uint32_t dup_lo(uint32_t v) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, 0x1010;" : "=r"(r) : "r"(v));
  return r;
}
__device__ __forceinline__ uint32_t dup_hi(uint32_t v) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %1, 0x3232;" : "=r"(r) : "r"(v));
  return r;
}

template <int S2_TWO>
__global__ void __launch_bounds__(256, 3)
dwconv7_s2m2_tiled_kernel(const __nv_bfloat16* __restrict__ in, const float* __restrict__ w,
                          const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W,
                          int Cin, int act) {
  constexpr int K = 7;
  using G2 = S2Geo<S2_TWO>;
  constexpr int S2_THO = G2::THO, S2_IW = G2::IW, S2_IH = G2::IH, S2_XP = G2::XP, S2_IN_BYTES = G2::IN_BYTES;
  extern __shared__ __align__(16) uint8_t smem_dw2[];
  const uint32_t s_in = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dw2));
  uint32_t* s_wh = reinterpret_cast<uint32_t*>(smem_dw2 + S2_IN_BYTES);  // [49][16] half2 = out-channel pairs
  const int Ho = H / 2, Wo = W / 2, Cout = 2 * Cin;
  const int tiles_x = Wo / S2_TWO;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int co0 = blockIdx.y * S2_CBO, ci0 = co0 / 2;
  const int b = blockIdx.z;
  const int xo0 = tx * S2_TWO, yo0 = ty * S2_THO;
  const int tid = threadIdx.x;

  // ---- stage the halo tile: 16 input channels per pixel = two 16-byte loads, bf16 -> fp16 ----
  const __nv_bfloat16* img = in + static_cast<size_t>(b) * H * W * Cin + ci0;
  constexpr int NV = S2_IH * S2_IW * 2;
  for (int idx = tid; idx < NV; idx += 256) {
    const int hv = idx & 1;  // which 8 of the 16 input channels
    const int xi = (idx >> 1) % S2_IW;
    const int r = (idx >> 1) / S2_IW;
    const int gy = 2 * yo0 + r - 3, gx = 2 * xo0 + xi - 3;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (gy >= 0 && gy < H && gx >= 0 && gx < W)
      v = __ldg(reinterpret_cast<const uint4*>(img + (static_cast<size_t>(gy) * W + gx) * Cin + hv * 8));
    const uint32_t slot = static_cast<uint32_t>(xi + (xi >> 3)) * 8u;
    const uint32_t d0 = s_in + static_cast<uint32_t>((r * 4 + 2 * hv) * S2_XP) * 8u + slot;
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d0), "r"(bf16x2_to_f16x2_sat(v.x)),
                 "r"(bf16x2_to_f16x2_sat(v.y))
                 : "memory");
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(d0 + S2_XP * 8u), "r"(bf16x2_to_f16x2_sat(v.z)),
                 "r"(bf16x2_to_f16x2_sat(v.w))
                 : "memory");
  }
  for (int idx = tid; idx < K * K * (S2_CBO / 2); idx += 256) {
    const int t = idx / (S2_CBO / 2), cp = idx % (S2_CBO / 2);
    const float w0 = __ldg(w + static_cast<size_t>(t) * Cout + co0 + 2 * cp);
    const float w1 = __ldg(w + static_cast<size_t>(t) * Cout + co0 + 2 * cp + 1);
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(w1), "f"(w0));
    s_wh[idx] = r;
  }
  __syncthreads();

  // ---- compute: warp = output row, lane = (vector of 4 input = 8 output channels, group of 4 output px) ----
  constexpr int PGS = S2_TWO / 4;                       // pixel groups of 4 per output row
  const int lane = tid & 31;
  const int cv = lane & 3, pg = (lane >> 2) % PGS;
  const int row = (tid >> 5) * (8 / PGS) + (lane >> 2) / PGS;
  float acc[4][8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + co0 + cv * 8));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + co0 + cv * 8) + 1);
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
      acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
    }
  }
  uint32_t racc[4][4];
#pragma unroll 1
  for (int ky = 0; ky < K; ++ky) {
    uint32_t wr[K][4];
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const uint4 wv = *reinterpret_cast<const uint4*>(s_wh + (ky * K + kx) * (S2_CBO / 2) + cv * 4);
      wr[kx][0] = wv.x; wr[kx][1] = wv.y; wr[kx][2] = wv.z; wr[kx][3] = wv.w;
    }
    const bool fresh = (ky & 1) == 0;
    const uint32_t line = s_in + static_cast<uint32_t>(((2 * row + ky) * 4 + cv) * S2_XP) * 8u;
#pragma unroll
    for (int i = 0; i < 2 * 3 + K; ++i) {  // inputs 2o + kx, o < 4, kx < 7
      const int xi = pg * 8 + i;
      uint32_t a, bb;
      asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(bb) : "r"(line + static_cast<uint32_t>(xi + (xi >> 3)) * 8u));
      const uint32_t x[4] = {dup_lo(a), dup_hi(a), dup_lo(bb), dup_hi(bb)};
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        const int kx = i - 2 * o;
        if (kx < 0 || kx >= K) continue;
        if (kx == 0 && fresh) {
#pragma unroll
          for (int c = 0; c < 4; ++c) racc[o][c] = hmul2_u(x[c], wr[0][c]);
        } else {
#pragma unroll
          for (int c = 0; c < 4; ++c) racc[o][c] = hfma2_u(x[c], wr[kx][c], racc[o][c]);
        }
      }
    }
    if (!fresh || ky == K - 1) {
#pragma unroll
      for (int o = 0; o < 4; ++o)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float2 f = h2_to_f2(racc[o][c]);
          acc[o][2 * c] += f.x;
          acc[o][2 * c + 1] += f.y;
        }
    }
  }
  __nv_bfloat16* orow =
      out + ((static_cast<size_t>(b) * Ho + (yo0 + row)) * Wo + xo0 + pg * 4) * Cout + co0 + cv * 8;
#pragma unroll
  for (int o = 0; o < 4; ++o) {
    Vec8<__nv_bfloat16> r;
#pragma unroll
    for (int c = 0; c < 8; ++c) r.v[c] = acc[o][c];
    if (act == ACT_GELU) {
#pragma unroll
      for (int c = 0; c < 8; ++c) r.v[c] = gelu_tanh_fit(r.v[c]);
    }
    r.store(orow + static_cast<size_t>(o) * Cout);
  }
}

}  // namespace

bool dwconv_tiled_supported(int dtype, int H, int W, int C, int mult, int k, int stride) {
  return dtype == DT_BF16 && stride == 1 && mult == 1 && (k == 3 || k == 7) && W % 64 == 0 &&
         H % TH == 0 && C % CB == 0;
}

namespace {
bool s2_fits(int H, int W, int two) { return (W / 2) % two == 0 && (H / 2) % (256 / two) == 0; }
}  // namespace

bool dwconv_s2m2_tiled_supported(int dtype, int H, int W, int Cin, int mult, int k, int stride) {
  return dtype == DT_BF16 && stride == 2 && mult == 2 && k == 7 && H % 2 == 0 && W % 2 == 0 &&
         (s2_fits(H, W, 32) || s2_fits(H, W, 16)) && (2 * Cin) % S2_CBO == 0;
}

int dwconv_s2m2_tiled(const void* in, const float* w, const float* bias, void* out, int B, int H, int W, int Cin,
                      int act, cudaStream_t stream) {
  if (s2_fits(H, W, 32)) {
    using G2 = S2Geo<32>;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(dwconv7_s2m2_tiled_kernel<32>), G2::SMEM)) return rc;
    dim3 grid((W / 2 / 32) * (H / 2 / G2::THO), 2 * Cin / S2_CBO, B);
    dwconv7_s2m2_tiled_kernel<32><<<grid, 256, G2::SMEM, stream>>>(static_cast<const __nv_bfloat16*>(in), w, bias,
                                                                    static_cast<__nv_bfloat16*>(out), H, W, Cin, act);
  } else {
    using G2 = S2Geo<16>;
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(dwconv7_s2m2_tiled_kernel<16>), G2::SMEM)) return rc;
    dim3 grid((W / 2 / 16) * (H / 2 / G2::THO), 2 * Cin / S2_CBO, B);
    dwconv7_s2m2_tiled_kernel<16><<<grid, 256, G2::SMEM, stream>>>(static_cast<const __nv_bfloat16*>(in), w, bias,
                                                                    static_cast<__nv_bfloat16*>(out), H, W, Cin, act);
  }
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int dwconv_tiled(const void* in, const float* w, const float* bias, void* out, int B, int H, int W,
                 int C, int k, int act, cudaStream_t stream) {
  if (k == 7) return launch_tiled_h<7>(in, w, bias, out, B, H, W, C, act, stream);
  return launch_tiled_h<3>(in, w, bias, out, B, H, W, C, act, stream);
}

}  // namespace fvla
