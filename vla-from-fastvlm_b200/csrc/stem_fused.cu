// FastViTHD stem.0 + stem.1 fused (bf16 mode):
//   y = GELU(dwConv3x3 s2 (GELU(Conv3x3 s2 (x, 3 -> C) + b0)) + b1)        [EXT FastViTHD patch_embed.0/.1]
// Unfused, stem.0's output (B x 512^2 x 96 bf16 = 1.6 GB at B = 32) is written by one kernel and read back by the
// next: 3.2 GB of HBM traffic per 32 images for a tensor that exists only to be pooled 2x2.  Here a CTA produces an
// 8 x 8 tile of the FINAL 256^2 map for 32 channels and keeps the 17 x 17 patch of the intermediate map it needs
// in shared memory (fp16):
//   1. stage the 35 x 35 input patch (NHWC4 bf16, zero outside the image = conv padding);
//   2. stem.0 on the tensor cores: per 16 intermediate pixels an implicit-GEMM with K = 9 taps x 4 channels
//      (4th channel zero) padded to 48: A fragments are gathered straight from the input patch (a channel pair of
//      one input pixel is one 32-bit shared load), B fragments (this CTA's 32 output channels) live in registers;
//      epilogue + b0, GELU on packed half pairs, zero outside the intermediate map (= stem.1's padding);
//   3. stem.1: depthwise 3x3 stride 2 with packed-half FMAs flushed to fp32, + b1, GELU, bf16 store.
// Algorithmic HBM traffic: read the ingested image once per channel block (L2 serves the repeats) + write the
// 256^2 map: ~0.8 GB per 32 images instead of 4.3 GB (im2col + GEMM + depthwise).
#include <cuda.h>

#include "common.cuh"
#include "epilogue_math.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tma_host.h"

#include <algorithm>
#include <cstring>

namespace fvla {
namespace {
using namespace epi;

constexpr int SF_TH = 8, SF_TW = 8;             // final-map tile (small: occupancy, not reuse, sets the speed)
constexpr int SF_CH = 2 * SF_TH + 1;            // 17 intermediate rows
constexpr int SF_CW = 2 * SF_TW + 1;            // 33 intermediate columns
constexpr int SF_NPX = SF_CH * SF_CW;           // 561 intermediate pixels
constexpr int SF_IH = 2 * SF_CH + 1;            // 35 input rows
constexpr int SF_IW = 2 * SF_CW + 1;            // 35 input columns
constexpr int SF_IP = SF_IW + 1;                // patch pitch: the TMA box starts one pixel early (16-byte alignment)
constexpr int SF_CB = 32;                       // output channels per CTA
constexpr int SF_CPITCH = 80;                   // bytes per intermediate pixel (32 fp16 + pad: conflict-free stores)
constexpr int SF_IN_BYTES = SF_IH * SF_IP * 8;     // 10 080 B landed by one TMA load
constexpr int SF_C_BYTES = SF_NPX * SF_CPITCH;
constexpr int SF_W1_BYTES = 9 * (SF_CB / 2) * 4;
constexpr int SF_BT_BYTES = 3 * 4 * 32 * 8;       // this channel block's B fragments
constexpr int SF_IN_SLOT = ((SF_IN_BYTES + 127) / 128) * 128;   // two patch buffers: the next tile's TMA load is in flight
constexpr int SF_SMEM = 2 * SF_IN_SLOT + SF_C_BYTES + SF_W1_BYTES + SF_BT_BYTES + 16 + 128;
constexpr int SF_BTAB_WORDS = 3 * 4 * 32 * 2;   // per channel block: [k-step 3][n-block 4][lane 32][2]

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t hfma2_s(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t d;
  asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ uint32_t hmul2_s(uint32_t a, uint32_t b) {
  uint32_t d;
  asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
  return d;
}
__device__ __forceinline__ float2 h2f2(uint32_t h) {
  float2 f;
  asm("{\n\t.reg .b16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}\n"
      : "=f"(f.x), "=f"(f.y)
      : "r"(h));
  return f;
}

// PERSISTENT: a CTA keeps one channel block (its stem.0 B fragments, stem.1 taps and biases are staged once) and walks
// the (image, tile) list with stride gridDim.x; the input patch of the NEXT tile is requested before this tile's
// compute starts (two patch buffers, one mbarrier each), so the TMA round trip and the per-CTA prologue — which made up
// a third of a one-tile CTA's life (ncu: 27 % issue utilisation at 4 CTAs per SM) — no longer sit on the critical path.
__global__ void __launch_bounds__(256, 4)
stem_fused_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint32_t* __restrict__ btab,
                  const float* __restrict__ b0_half, const float* __restrict__ w1, const float* __restrict__ b1,
                  __nv_bfloat16* __restrict__ out, int S, int C, int tiles_per_img, int total_tiles) {
  extern __shared__ __align__(16) uint8_t smem_raw_sf[];
  uint8_t* smem_sf = smem_raw_sf + ((128u - (ptx::smem_u32(smem_raw_sf) & 127u)) & 127u);  // TMA destination alignment
  const uint32_t s_in0 = ptx::smem_u32(smem_sf);
  const uint32_t s_c = s_in0 + 2 * SF_IN_SLOT;
  uint32_t* s_w1 = reinterpret_cast<uint32_t*>(smem_sf + 2 * SF_IN_SLOT + SF_C_BYTES);  // [9][16] half2
  const uint32_t s_bar = s_c + SF_C_BYTES + SF_W1_BYTES + SF_BT_BYTES;  // two barriers, 8 B each

  const int So = S / 4, Sc = S / 2;  // final / intermediate map side
  const int tiles_x = So / SF_TW;
  const int c0 = blockIdx.y * SF_CB;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  auto request = [&](int tile, int buf) {  // thread 0: TMA load of the input patch of `tile` into buffer `buf`
    const int bb = tile / tiles_per_img, r = tile - bb * tiles_per_img;
    const int tyy = r / tiles_x, txx = r - tyy * tiles_x;
    const int ix = 2 * (2 * txx * SF_TW - 1) - 1, iy = 2 * (2 * tyy * SF_TH - 1) - 1;
    ptx::mbar_arrive_expect_tx(s_bar + 8u * buf, SF_IN_BYTES);
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(s_in0 + static_cast<uint32_t>(buf) * SF_IN_SLOT), "l"(reinterpret_cast<uint64_t>(&tmap_in)),
        "r"((ix - 1) * 4), "r"(iy), "r"(bb), "r"(s_bar + 8u * buf)
        : "memory");
  };

  // ---- 1. input patch (4 bf16 per pixel = 8 bytes): ONE TMA load of 35 rows x 36 pixels, zero fill outside the
  // image (= conv padding).  The image is addressed as [B][S][4 S] so a row of pixels is one contiguous TMA row; the
  // box starts at the even pixel ix0 - 1 (16-byte aligned), patch pixel (r, x) sits at (r * 36 + x + 1) * 8.
  // (The per-pixel load loop this replaces was 12 % of the kernel's instructions and 21 % of its stall samples.)
  if (tid == 0) {
    ptx::mbar_init(s_bar, 1);
    ptx::mbar_init(s_bar + 8u, 1);
    ptx::fence_barrier_init();
  }
  for (int idx = tid; idx < 9 * (SF_CB / 2); idx += 256) {
    const int tap = idx / (SF_CB / 2), cp = idx % (SF_CB / 2);
    const float wa = __ldg(w1 + static_cast<size_t>(tap) * C + c0 + 2 * cp);
    const float wb = __ldg(w1 + static_cast<size_t>(tap) * C + c0 + 2 * cp + 1);
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(wb), "f"(wa));
    s_w1[idx] = r;
  }
  // B fragments of stem.0 for this channel block (3 k-steps x 4 n-blocks x 32 lanes x 2 words) -> shared memory:
  // 24 registers per thread less, which is what lets four CTAs share an SM
  const uint32_t s_bt = s_c + SF_C_BYTES + SF_W1_BYTES;
  for (int idx = tid; idx < SF_BTAB_WORDS / 2; idx += 256) {
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(btab + static_cast<size_t>(blockIdx.y) * SF_BTAB_WORDS) + idx);
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(s_bt + static_cast<uint32_t>(idx) * 8u), "r"(v.x), "r"(v.y)
                 : "memory");
  }
  float bias0[4][2];
#pragma unroll
  for (int nb = 0; nb < 4; ++nb) {
    bias0[nb][0] = __ldg(b0_half + c0 + nb * 8 + 2 * t);
    bias0[nb][1] = __ldg(b0_half + c0 + nb * 8 + 2 * t + 1);
  }
  __syncthreads();
  pdl_sync();  // the weight staging above overlaps the previous kernel; the frames are its output
  if (tid == 0 && static_cast<int>(blockIdx.x) < total_tiles) request(blockIdx.x, 0);

  int it = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
  const int buf = it & 1;
  const uint32_t s_in = s_in0 + static_cast<uint32_t>(buf) * SF_IN_SLOT;
  const int b = tile / tiles_per_img, trem = tile - b * tiles_per_img;
  const int ty = trem / tiles_x, tx = trem - ty * tiles_x;
  const int xo0 = tx * SF_TW, yo0 = ty * SF_TH;
  const int cy0 = 2 * yo0 - 1, cx0 = 2 * xo0 - 1;   // intermediate-map origin of the patch
  // the other patch buffer was last read in the previous iteration's phase 2, which every warp has left (the
  // __syncthreads between the phases below): request the next tile into it now
  if (tid == 0 && tile + static_cast<int>(gridDim.x) < total_tiles) {
    ptx::fence_proxy_async_smem();  // generic-proxy reads of that buffer (stem.0's gathers) before the TMA rewrites it
    request(tile + gridDim.x, buf ^ 1);
  }
  ptx::mbar_wait(s_bar + 8u * buf, static_cast<uint32_t>(it >> 1) & 1u);  // this tile's input patch has landed

  // ---- 2. stem.0: 16 intermediate pixels per mma tile, pixels flattened over the 17 x 33 patch ----
  // k = 4 * tap + ci (tap = ky * 3 + kx); this lane's k pairs: taps (4s + t/2) and (4s + 2 + t/2), channels 2(t&1)..+1
  const int tap_sub = t >> 1, cpair = t & 1;
  for (int mt = warp; mt < (SF_NPX + 15) / 16; mt += 8) {
    uint32_t a[3][4];
    int pbase[2];
    bool pvalid[2], pinside[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int p = mt * 16 + g + 8 * h;
      pvalid[h] = p < SF_NPX;
      const int cy = p / SF_CW, cx = p - cy * SF_CW;
      // input pixel of tap (0,0): row 2*cy (patch-local, since iy0 = 2*cy0 - 1), column 2*cx
      pbase[h] = pvalid[h] ? ((2 * cy) * SF_IP + 2 * cx + 1) * 8 + cpair * 4 : 0;
      const int gcy = cy0 + cy, gcx = cx0 + cx;
      pinside[h] = pvalid[h] && gcy >= 0 && gcy < Sc && gcx >= 0 && gcx < Sc;
    }
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int kh = 0; kh < 2; ++kh) {
        const int tap = 4 * s + 2 * kh + tap_sub;
        const int ky = tap / 3, kx = tap - ky * 3;
        const uint32_t off = static_cast<uint32_t>((ky * SF_IP + kx) * 8);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v = 0u;
          if (tap < 9 && pvalid[h])
            asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(s_in + static_cast<uint32_t>(pbase[h]) + off));
          a[s][2 * kh + h] = v;  // a0: row g k-lo, a1: row g+8 k-lo, a2: row g k-hi, a3: row g+8 k-hi
        }
      }
    float acc[4][4];
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
      acc[nb][0] = acc[nb][2] = bias0[nb][0];
      acc[nb][1] = acc[nb][3] = bias0[nb][1];
    }
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        uint32_t q0, q1;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];"
                     : "=r"(q0), "=r"(q1)
                     : "r"(s_bt + static_cast<uint32_t>(((s * 4 + nb) * 32 + lane) * 8)));
        mma16816(acc[nb], a[s], q0, q1);
      }
    // GELU (operands are x/2: weights and bias pre-halved) on half pairs; zero where stem.1 sees padding
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (!pvalid[h]) continue;
      const int p = mt * 16 + g + 8 * h;
      const uint32_t dst = s_c + static_cast<uint32_t>(p) * SF_CPITCH + static_cast<uint32_t>(t) * 4u;
#pragma unroll
      for (int nb = 0; nb < 4; ++nb) {
        const uint32_t v = pinside[h] ? gelu_half_f16x2(acc[nb][2 * h], acc[nb][2 * h + 1]) : 0u;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(dst + nb * 16), "r"(v) : "memory");
      }
    }
  }
  __syncthreads();

  // ---- 3. stem.1: depthwise 3x3 stride 2 over the fp16 patch; thread = 8 channels x 2 adjacent outputs ----
  {
    const int cv = tid & 3, pg = tid >> 2;           // pixel pairs
    const int row = pg / (SF_TW / 2), xo = (pg % (SF_TW / 2)) * 2;  // output row, first output column of the pair
    if (pg < SF_TH * SF_TW / 2) {
    float acc[2][8];
    {
      const float4 q0 = __ldg(reinterpret_cast<const float4*>(b1 + c0 + cv * 8));
      const float4 q1 = __ldg(reinterpret_cast<const float4*>(b1 + c0 + cv * 8) + 1);
#pragma unroll
      for (int o = 0; o < 2; ++o) {
        acc[o][0] = q0.x; acc[o][1] = q0.y; acc[o][2] = q0.z; acc[o][3] = q0.w;
        acc[o][4] = q1.x; acc[o][5] = q1.y; acc[o][6] = q1.z; acc[o][7] = q1.w;
      }
    }
    uint32_t racc[2][4];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      uint32_t wr[3][4];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const uint4 wv = *reinterpret_cast<const uint4*>(s_w1 + (ky * 3 + kx) * (SF_CB / 2) + cv * 4);
        wr[kx][0] = wv.x; wr[kx][1] = wv.y; wr[kx][2] = wv.z; wr[kx][3] = wv.w;
      }
      const bool fresh = ky != 1;  // rows 0 and 2 start a partial sum, row 1 continues row 0's
      const uint32_t line = s_c + static_cast<uint32_t>(((2 * row + ky) * SF_CW + 2 * xo) * SF_CPITCH + cv * 16);
#pragma unroll
      for (int i = 0; i < 5; ++i) {
        uint32_t x[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3])
                     : "r"(line + static_cast<uint32_t>(i * SF_CPITCH)));
#pragma unroll
        for (int o = 0; o < 2; ++o) {
          const int kx = i - 2 * o;
          if (kx < 0 || kx > 2) continue;
          if (kx == 0 && fresh) {
#pragma unroll
            for (int c = 0; c < 4; ++c) racc[o][c] = hmul2_s(x[c], wr[0][c]);
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) racc[o][c] = hfma2_s(x[c], wr[kx][c], racc[o][c]);
          }
        }
      }
      if (ky >= 1) {  // flush rows (0,1) together, then row 2
#pragma unroll
        for (int o = 0; o < 2; ++o)
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const float2 f = h2f2(racc[o][c]);
            acc[o][2 * c] += f.x;
            acc[o][2 * c + 1] += f.y;
          }
      }
    }
    __nv_bfloat16* orow = out + ((static_cast<size_t>(b) * So + yo0 + row) * So + xo0 + xo) * C + c0 + cv * 8;
#pragma unroll
    for (int o = 0; o < 2; ++o) {
      Vec8<__nv_bfloat16> r;
#pragma unroll
      for (int c = 0; c < 8; ++c) r.v[c] = gelu_tanh_fit(acc[o][c]);
      r.store(orow + static_cast<size_t>(o) * C);
    }
    }
  }
  __syncthreads();  // the intermediate patch is rewritten by the next tile's stem.0
  }
}

}  // namespace

size_t stem_fused_btab_words(int C) { return static_cast<size_t>(C / SF_CB) * SF_BTAB_WORDS; }

// w0_packed: [27][C] fp32 with row (ky*3+kx)*3 + ci (the layout stem_conv3x3_s2 takes).  Table word order:
// [channel block][k-step][n-block][lane][2]; B[k][n] with k = 4*tap + ci, pre-halved for the GELU-from-x/2 form.
void stem_fused_build_btab(const float* w0_packed, int C, uint32_t* btab_host) {
  auto bf16_bits = [](float f) -> uint32_t {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t r = u + 0x7fffu + ((u >> 16) & 1u);  // round to nearest even
    return r >> 16;
  };
  auto wk = [&](int k, int ch) -> float {  // element of the [48][C] implicit-GEMM matrix
    const int tap = k >> 2, ci = k & 3;
    if (tap >= 9 || ci >= 3) return 0.0f;
    return 0.5f * w0_packed[static_cast<size_t>(tap * 3 + ci) * C + ch];
  };
  for (int cb = 0; cb < C / SF_CB; ++cb)
    for (int s = 0; s < 3; ++s)
      for (int nb = 0; nb < 4; ++nb)
        for (int lane = 0; lane < 32; ++lane) {
          const int g = lane >> 2, t = lane & 3;
          const int ch = cb * SF_CB + nb * 8 + g;
          const int k0 = 16 * s + 2 * t;
          uint32_t* dst = btab_host + ((static_cast<size_t>(cb) * 3 + s) * 4 + nb) * 64 + lane * 2;
          dst[0] = bf16_bits(wk(k0, ch)) | (bf16_bits(wk(k0 + 1, ch)) << 16);
          dst[1] = bf16_bits(wk(k0 + 8, ch)) | (bf16_bits(wk(k0 + 9, ch)) << 16);
        }
}

bool stem_fused_supported(int dtype, int S, int C) {
  return dtype == DT_BF16 && C % SF_CB == 0 && S % 4 == 0 && (S / 4) % SF_TW == 0 && (S / 4) % SF_TH == 0;
}

int stem_fused(const void* in, const uint32_t* btab, const float* b0_half, const float* w1_packed, const float* b1,
               void* out, int B, int S, int C, cudaStream_t stream) {
  FVLA_REQUIRE(stem_fused_supported(DT_BF16, S, C), "stem_fused: unsupported geometry");
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(stem_fused_kernel), SF_SMEM)) return rc;
  // the ingested image [B][S][S][4] bf16 as a 3-D tensor [B][S][4 S]: box = 36 pixels x 35 rows
  TmaEncodeTiledFn fn = tma_encode_fn();
  FVLA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FVLA_REQUIRE((reinterpret_cast<uintptr_t>(in) & 15u) == 0, "TMA base must be 16-byte aligned");
  CUtensorMap ti;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(S) * 4, static_cast<cuuint64_t>(S), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(S) * 8, static_cast<cuuint64_t>(S) * S * 8};
  cuuint32_t box[3] = {SF_IP * 4, SF_IH, 1u};
  cuuint32_t estr[3] = {1u, 1u, 1u};
  CUresult r = fn(&ti, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(in), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (stem) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  const int So = S / 4;
  const int tiles_per_img = (So / SF_TW) * (So / SF_TH);
  const long long total = static_cast<long long>(tiles_per_img) * B;
  FVLA_REQUIRE(total < (1ll << 30), "stem_fused: too many tiles");
  // persistent CTAs: 4 per SM share the 3 channel blocks, each walking the tile list with a fixed channel block
  const int per_cb = std::max(1, (4 * num_sms()) / (C / SF_CB));
  dim3 grid(static_cast<unsigned>(total < per_cb ? total : per_cb), C / SF_CB, 1);
  FVLA_CUDA_CHECK(launch_pdl(stem_fused_kernel, grid, dim3(256), SF_SMEM, stream, ti, btab, b0_half, w1_packed, b1,
                             static_cast<__nv_bfloat16*>(out), S, C, tiles_per_img, static_cast<int>(total)));
  return 0;
}

}  // namespace fvla
