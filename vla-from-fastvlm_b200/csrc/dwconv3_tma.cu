// Depthwise 3x3 (stride 1, NHWC bf16, RepMixer token mixer of the wide FastViTHD stages) as a persistent,
// TMA-fed kernel with fp32 accumulation on packed FFMA2.
//
// Why a second 3x3 kernel: ncu of the register-staged tile kernel (dwconv_tiled.cu) at 128^2 x 192 showed it
// 64 % issue-bound at 2.9 TB/s — 853 instructions per thread and tile, of which only 144 are the HFMA2 taps:
// the rest is index arithmetic of the staged loads, the bf16 -> fp16 conversion (3 slots per value pair) and
// the fp16 -> fp32 flushes (4 slots per pair, twice per tile).  Here
//   * the (8+2) x (32+2) x 32-channel halo tile arrives by ONE 4-D TMA load per tile (zero fill outside the
//     image, no per-element index arithmetic, no staging registers), four tiles deep, so HBM requests stay in
//     flight across tile boundaries;
//   * the taps run on fma.rn.f32x2 over register pairs (two channels per issue slot, like HFMA2) straight into
//     fp32 accumulators: unpacking a bf16 pair costs two ALU slots (shift / mask) and there is nothing to flush —
//     ~430 instructions per thread and tile, and the result is the exact fp32 sum rounded once to bf16;
//   * a CTA walks tiles with the channel block as the fastest index, so CTAs resident together consume whole
//     128-byte lines of the NHWC tensor.
// Measured at 128^2 x 192, batch 32: 0.118 -> 0.090 ms (3.4 -> 4.5 TB/s); 46 M instead of 84 M warp instructions.
// A variant without the stores runs in 0.084 ms and one without the taps in 0.079 ms: the load -> unpack -> FFMA2
// chain (16 warps per SM, latency-bound at ~50 % issue utilisation; FFMA2 holds the FMA pipe for two cycles) and
// the memory stream each need ~90 % of the time, so what is left is their overlap.  A dedicated producer warp and a
// per-tile __syncthreads version measured the same.
// Shared-memory layout of a tile: [row][pixel][32 channels] bf16, 64 B per pixel, dense (no swizzle: the reads below
// are conflict-free without one, and every load is base register + immediate).  The box is 35 pixels wide (halo +
// one spare), so the row pitch is an odd multiple of 64 B and consecutive rows start in opposite halves of the
// 128-byte bank row.  A warp owns two output rows x 16 pixels; lane = (channel vector cv, row of the pair, group of
// 4 pixels): the eight 16-byte reads of one LDS.128 wavefront (4 cv x 2 rows, same pixel) land in eight distinct
// bank groups.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "epilogue_math.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tma_host.h"

namespace fvla {
namespace {

constexpr int TH = 8;                 // output rows per tile
constexpr int CONSUMER_WARPS = 8;     // one-row kernel; the two-row kernel runs 4 warps (Geo::WARPS)
constexpr int NSTAGE = 4;

// CB = channels per tile.  8 x 32 pixels x 32 channels is the shape the engine's layers use; the 8 x 16 x 64 shape
// (128-byte TMA rows, one pixel = one bank row) measured 10 % slower at 128^2 x 192 and only serves widths that are a
// multiple of 16 but not of 32.
// R2 (CB = 32 only): TWO output rows per thread, 128-thread CTAs, three per SM.  ncu of the one-row kernel
// (profiles/r02_ncu_full_dwconv3_tma_c192.txt): the LSU data pipe is 76 % busy and half of its shared-memory wavefronts
// are the per-tile TAP loads — every quarter-warp re-reads the same 128 bytes, 20 LDS.128 per lane for 32 outputs.  With
// two rows per thread the taps serve 64 outputs (56 instead of 76 LDS.128 per 64 outputs).  A thread's rows are two apart
// (r, r+2) so that the two threads sharing a quarter-warp wavefront sit on ADJACENT rows, whose pitch is an odd
// multiple of 64 B: conflict-free like the one-row mapping.
template <int CB_, bool R2_ = false> struct Geo {
  static constexpr int CB = CB_;
  static constexpr bool R2 = R2_;
  static constexpr int WARPS = R2 ? 4 : CONSUMER_WARPS;
  static constexpr int NTHREADS = WARPS * 32;
  static constexpr int STAGES = R2 ? 3 : NSTAGE;
  static constexpr int CTAS = R2 ? 3 : 2;
  static constexpr int NCV = CB / 8;                       // 16-byte channel vectors per pixel
  static constexpr int TW = CB == 32 ? 32 : 16;            // output pixels per tile row
  // CB 32: 64 B per pixel; one spare column makes the row pitch an ODD multiple of 64 B, so consecutive rows start
  // in opposite halves of the 128-byte bank row.  CB 64: a pixel is a whole bank row, nothing to skew.
  static constexpr int IW = TW + (CB == 32 ? 3 : 2);
  static constexpr int IH = TH + 2;
  static constexpr int PXB = CB * 2;                       // bytes per pixel
  static constexpr int TILE_B = IH * IW * PXB;             // bytes landed per tile (22 400 / 23 040)
  static constexpr int SLOT_B = (TILE_B + 127) / 128 * 128;
  static constexpr int WSLOT_B = 10 * CB * 4;              // 9 tap rows + bias, fp32
  static constexpr int STAGE_B = SLOT_B + WSLOT_B;         // halo tile + its channel block's taps and bias
  static constexpr int SMEM_B = STAGES * STAGE_B + 2 * STAGES * 8 + 1024;
};

using epi::f32x2;
using epi::fma2;
using epi::pk2;
using epi::upk2;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
      "[%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}

// Tile ids are mixed-radix numbers (image, tile row, tile column, channel block) with the channel block as the
// lowest digit.  A CTA visits id = blockIdx.x + n * gridDim.x: the digits are decoded once (integer divisions) and
// then advanced by the decoded stride with carries, which costs a dozen slots per tile instead of three divisions.
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

struct TileIt {
  int cb, tx, ty, b;
};
__device__ __forceinline__ TileIt decode_tile(int id, int n_cblk, int tiles_x, int tiles_y) {
  TileIt t;
  t.cb = id % n_cblk;
  id /= n_cblk;
  t.tx = id % tiles_x;
  id /= tiles_x;
  t.ty = id % tiles_y;
  t.b = id / tiles_y;
  return t;
}
__device__ __forceinline__ void advance(TileIt& t, const TileIt& d, int n_cblk, int tiles_x, int tiles_y) {
  t.cb += d.cb;
  if (t.cb >= n_cblk) { t.cb -= n_cblk; ++t.tx; }
  t.tx += d.tx;
  if (t.tx >= tiles_x) { t.tx -= tiles_x; ++t.ty; }
  t.ty += d.ty;
  if (t.ty >= tiles_y) { t.ty -= tiles_y; ++t.b; }
  t.b += d.b;
}

// Per tile one 4-D TMA load of the halo tile plus ten 128-byte bulk copies (9 tap rows + bias of the channel
// block) complete on the slot's `full` barrier; each warp releases the slot on its `empty` barrier once it has read
// its inputs, and thread 0 re-arms the slot one tile later.  Warps drift apart by up to NSTAGE - 1 tiles and there is
// no CTA-wide barrier in the loop (ncu of the first version: 17 % of samples at a per-tile __syncthreads and 8 %
// waiting for the weight LDG before its STS).
template <int CB, bool R2>
__global__ void __launch_bounds__(Geo<CB, R2>::NTHREADS, Geo<CB, R2>::CTAS)
dwconv3_tma_kernel(const __grid_constant__ CUtensorMap tmap_in, const float* __restrict__ w,
                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int C,
                   int tiles_x, int tiles_y, int n_cblk, int total_tiles) {
  using G = Geo<CB, R2>;
  constexpr int TW = G::TW, IW = G::IW, STAGE_B = G::STAGE_B, SLOT_B = G::SLOT_B, NSTAGE = G::STAGES;
  constexpr int NR = R2 ? 2 : 1;   // output rows per thread
  extern __shared__ uint8_t smem_dw3[];
  const uint32_t base = (ptx::smem_u32(smem_dw3) + 1023u) & ~1023u;
  const uint8_t* gbase = smem_dw3 + (base - ptx::smem_u32(smem_dw3));
  const uint32_t bars = base + NSTAGE * STAGE_B;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  auto empty_bar = [&](int s) { return bars + 8u * (NSTAGE + s); };
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_in);
    for (int s = 0; s < NSTAGE; ++s) {
      ptx::mbar_init(full_bar(s), 1);
      ptx::mbar_init(empty_bar(s), G::WARPS);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  pdl_sync();

  const int first = blockIdx.x, stride = gridDim.x;
  const TileIt step = decode_tile(stride, n_cblk, tiles_x, tiles_y);
  TileIt cur = decode_tile(first, n_cblk, tiles_x, tiles_y);
  TileIt pre = cur;  // thread 0: the next tile to request
  auto issue_load = [&](int s) {
    const uint32_t dst = base + s * STAGE_B;
    const int c0 = pre.cb * CB;
    ptx::mbar_arrive_expect_tx(full_bar(s), G::TILE_B + G::WSLOT_B);
    tma_load_4d(dst, &tmap_in, c0, pre.tx * TW - 1, pre.ty * TH - 1, pre.b, full_bar(s));
#pragma unroll
    for (int t = 0; t < 9; ++t) bulk_g2s(dst + SLOT_B + t * CB * 4, w + static_cast<size_t>(t) * C + c0, CB * 4, full_bar(s));
    bulk_g2s(dst + SLOT_B + 9 * CB * 4, bias + c0, CB * 4, full_bar(s));
    advance(pre, step, n_cblk, tiles_x, tiles_y);
  };
  if (tid == 0)
    for (int s = 0; s < NSTAGE; ++s)
      if (first + s * stride < total_tiles) issue_load(s);

  // CB 32: warp = 2 rows x 16 pixels, lane = (cv 0..3, row of the pair, pixel group): lanes 4..7 of a quarter-warp
  // read the next row.  CB 64: warp = 1 row x 16 pixels, lane = (cv 0..7, pixel group).  Either way the eight
  // 16-byte reads of an LDS.128 wavefront cover one whole 128-byte bank row.
  // R2: warp = (4-row band, half of the 32 pixels), lane = (cv 0..3, row parity h, pixel group): rows band + h and band + h + 2.
  const int cv = lane % G::NCV;
  const int row = R2 ? (warp >> 1) * 4 + ((lane >> 2) & 1) : (CB == 32 ? (warp >> 1) * 2 + ((lane >> 2) & 1) : warp);
  const int px0 = CB == 32 ? (warp & 1) * 16 + (lane >> 3) * 4 : (lane >> 3) * 4;  // first of the lane's 4 output pixels
  // byte offset of this lane's first input (tile row `row`, pixel px0, channel vector cv) inside a slot
  const uint32_t lane_off = static_cast<uint32_t>((row * IW + px0) * G::PXB + cv * 16);

  int k = 0;
  for (int tile = first; tile < total_tiles; tile += stride, ++k) {
    const int s = k % NSTAGE;
    if (tid == 0 && k > 0 && tile + (NSTAGE - 1) * stride < total_tiles) {
      // refill the slot of the PREVIOUS tile: by now the other warps have normally released it, so thread 0
      // (which also computes) rarely waits here
      const int ps = (k - 1) % NSTAGE;
      ptx::mbar_wait(empty_bar(ps), static_cast<uint32_t>((k - 1) / NSTAGE) & 1u);
      issue_load(ps);
    }
    const TileIt tc = cur;
    advance(cur, step, n_cblk, tiles_x, tiles_y);
    const float* wt = reinterpret_cast<const float*>(gbase + s * STAGE_B + SLOT_B);
    const uint32_t src = base + s * STAGE_B + lane_off;
    ptx::mbar_wait(full_bar(s), static_cast<uint32_t>(k / NSTAGE) & 1u);

    f32x2 acc[NR * 4][4];
    {
      const float4 b0 = *reinterpret_cast<const float4*>(wt + 9 * CB + cv * 8);
      const float4 b1 = *reinterpret_cast<const float4*>(wt + 9 * CB + cv * 8 + 4);
#pragma unroll
      for (int o = 0; o < NR * 4; ++o) {
        acc[o][0] = pk2(b0.x, b0.y); acc[o][1] = pk2(b0.z, b0.w);
        acc[o][2] = pk2(b1.x, b1.y); acc[o][3] = pk2(b1.z, b1.w);
      }
    }
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      f32x2 wr[3][4];
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const float4 w0 = *reinterpret_cast<const float4*>(wt + (ky * 3 + kx) * CB + cv * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wt + (ky * 3 + kx) * CB + cv * 8 + 4);
        wr[kx][0] = pk2(w0.x, w0.y); wr[kx][1] = pk2(w0.z, w0.w);
        wr[kx][2] = pk2(w1.x, w1.y); wr[kx][3] = pk2(w1.z, w1.w);
      }
#pragma unroll
      for (int rr = 0; rr < NR; ++rr)
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        uint32_t x[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3])
                     : "r"(src + static_cast<uint32_t>(((ky + 2 * rr) * IW + i) * G::PXB)));
        f32x2 xf[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) xf[c] = pk2(__uint_as_float(x[c] << 16), __uint_as_float(x[c] & 0xffff0000u));
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          const int kx = i - o;
          if (kx < 0 || kx >= 3) continue;
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[rr * 4 + o][c] = fma2(xf[c], wr[kx][c], acc[rr * 4 + o][c]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(empty_bar(s));  // this warp has read everything it needs from slot s
    // ---- store: 4 pixels x 8 channels per lane, 16 bytes each ----
    __nv_bfloat16* orow = out + ((static_cast<size_t>(tc.b) * H + (tc.ty * TH + row)) * W + tc.tx * TW + px0) * C + tc.cb * CB + cv * 8;
#pragma unroll
    for (int rr = 0; rr < NR; ++rr)
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      uint4 v;
      float a, b2;
      upk2(acc[rr * 4 + o][0], a, b2); v.x = epi::pack_bf16(a, b2);
      upk2(acc[rr * 4 + o][1], a, b2); v.y = epi::pack_bf16(a, b2);
      upk2(acc[rr * 4 + o][2], a, b2); v.z = epi::pack_bf16(a, b2);
      upk2(acc[rr * 4 + o][3], a, b2); v.w = epi::pack_bf16(a, b2);
      *reinterpret_cast<uint4*>(orow + (static_cast<size_t>(2 * rr) * W + o) * C) = v;
    }
  }
}

template <int CB>
int make_tmap(CUtensorMap* out, const void* ptr, int B, int H, int W, int C) {
  using G = Geo<CB>;
  TmaEncodeTiledFn fn = tma_encode_fn();
  FVLA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FVLA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA base must be 16-byte aligned");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {G::CB, G::IW, G::IH, 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (dwconv3) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

template <int CB, bool R2>
int launch(const void* in, const float* w_packed, const float* bias, void* out, int B, int H, int W, int C,
           cudaStream_t stream) {
  using G = Geo<CB, R2>;
  auto kfn = dwconv3_tma_kernel<CB, R2>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), G::SMEM_B)) return rc;
  CUtensorMap ti;
  if (int rc = make_tmap<CB>(&ti, in, B, H, W, C)) return rc;
  const int tiles_x = W / G::TW, tiles_y = H / TH, n_cblk = C / CB;
  const long long total = static_cast<long long>(tiles_x) * tiles_y * n_cblk * B;
  FVLA_REQUIRE(total < (1ll << 31), "dwconv3_tma: too many tiles");
  const int resident = G::CTAS * num_sms();
  const int grid = total < resident ? static_cast<int>(total) : resident;
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(G::NTHREADS), G::SMEM_B, stream, ti, w_packed, bias,
                             static_cast<__nv_bfloat16*>(out), H, W, C, tiles_x, tiles_y, n_cblk,
                             static_cast<int>(total)));
  return 0;
}

}  // namespace

bool dwconv3_tma_supported(int dtype, int H, int W, int C, int mult, int k, int stride, int act) {
  if (!(dtype == DT_BF16 && k == 3 && stride == 1 && mult == 1 && act == ACT_NONE && H % TH == 0)) return false;
  return (C % 32 == 0 && W % Geo<32>::TW == 0) || (C % 64 == 0 && W % Geo<64>::TW == 0);
}

int dwconv3_tma(const void* in, const float* w_packed, const float* bias, void* out, int B, int H, int W, int C,
                cudaStream_t stream) {
  const bool r2 = std::getenv("FVLA_DWCONV3_ONE_ROW") == nullptr;  // A/B switch, read per call (tests toggle it)
  if (W % Geo<32>::TW == 0)
    return r2 ? launch<32, true>(in, w_packed, bias, out, B, H, W, C, stream)
              : launch<32, false>(in, w_packed, bias, out, B, H, W, C, stream);
  return launch<64, false>(in, w_packed, bias, out, B, H, W, C, stream);
}

}  // namespace fvla
