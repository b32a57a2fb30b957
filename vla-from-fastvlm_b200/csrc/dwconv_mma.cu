// Depthwise 7x7 convolution (stride 1, pad 3) on the tensor cores — the ConvFFN / RepCPE conv of every
// FastViTHD block (reference: HF-hub FastViTHD `convffn.conv` + folded BN).
//
// Why tensor cores for a depthwise conv: 49 MACs per output make the SIMT version issue-bound (ncu:
// profiles/r01_dwconv_k7_c192_ncu.txt — FMA+ALU pipes saturated at ~1 TB/s, a sixth of HBM speed).  Per
// channel the conv is   out[y][x] = sum_ky ( in[y+ky][x .. x+6] . w[ky][0..6] ),
// i.e. for every kernel row a product of a [rows x 16] slice of the image with a banded 16 x 8 Toeplitz
// matrix of that row's taps: one mma.sync.m16n8k16 per (kernel row, 16 rows x 8 outputs), 7 per 128 outputs
// instead of 3136 scalar FMAs.  44 % of the multiplies hit structural zeros; the tensor pipe has >20x the
// headroom, and the kernel becomes what it should be — bound by moving the feature map.
//
// The catch is layout: activations are NHWC (channel fastest) but an mma fragment wants, for ONE channel,
// rows = image rows and 16-bit pairs along x.  The kernel therefore transposes on chip, both ways, with
// matrix loads/stores instead of per-element shuffles:
//
//   TMA (4-D tensor map, SWIZZLE_64B, zero fill outside the image = the conv's zero padding)
//     -> slab ring   [2 rows][TW+8 px][32 ch]                      natural NHWC order
//   ldmatrix.x4.trans  (8 px x 8 ch tiles -> channel-major register pairs)  + 4 STS.32
//     -> planes      [32 ch][22 rows][TW+8 px]  bf16               one image plane per channel
//   ldmatrix.x4 (A fragments, rows interleaved even/odd so kernel row ky+1 reuses half of ky's tiles)
//     + Toeplitz B fragments from a 9-word-per-kernel-row table -> 7*NB mma.sync per (channel, 16 x 8NB px)
//   fp32 accumulators (+bias) -> bf16 -> output planes (over the channel's own input plane)
//     -> LDS.32 + stmatrix.x4.trans into a per-warp 8 px x 32 ch buffer -> 64-byte runs to global memory.
//
// CTA = 256 threads, tile = 16 rows x TW px x 32 channels, ~105 KB of shared memory -> 2 persistent CTAs per SM,
// each prefetching its next tile's first slabs while it computes.  Algorithmic traffic: read + write the map once
// (4 B per output element in bf16); the halo re-reads (22 x 40 / 16 x 32 = 1.7x) are served by L2 (ncu: DRAM
// bytes = algorithmic).  The kernel is bound by shared-memory bandwidth (~520 KB through smem per 64 KB of HBM
// traffic; budget in DESIGN.md 3.3), not by the tensor pipe (29 % busy) or HBM.
#include "common.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tma_host.h"

#include <cstdlib>

namespace fvla {
namespace {

constexpr int CB = 32;            // channels per CTA (64 B per pixel: one swizzle-64 row)
constexpr int TH = 16;            // output rows per CTA (one mma M tile)
constexpr int IH = TH + 6;        // input rows incl. halo
constexpr int NSLAB = IH / 2;     // TMA slabs of two rows
constexpr int RING = 6;           // slabs in flight
constexpr int WTAB_WORDS = 64;    // Toeplitz table words per channel (7 kernel rows x 9, padded)
constexpr int OROW_W = 18;        // output-plane row pitch in words (conflict-free D-fragment stores)
constexpr int OPLANE_W = 292;     // words an output plane needs (16 * 18 = 288, +4: = 4 (mod 8))
constexpr int THREADS = 256;

template <int NB> struct Geo {
  static constexpr int TW = 8 * NB;               // output pixels per tile row
  static constexpr int IW = TW + 8;               // staged input pixels per row (16-wide k chunks reach 8 past)
  static constexpr int XB = IW / 8;               // 8-pixel blocks per staged row
  static constexpr int ROW_B = IW * 2;            // plane row pitch (bytes): an odd number of 16-byte units
  static constexpr int PLANE_RAW_W = IH * ROW_B / 4;
  static constexpr int PLANE_IN_W = PLANE_RAW_W + ((4 - (PLANE_RAW_W & 7)) & 7);  // = 4 (mod 8) words
  // a channel's output plane overwrites its own input plane (only one warp ever reads it), so one pitch
  static constexpr int PLANE_W = PLANE_IN_W > OPLANE_W ? PLANE_IN_W : OPLANE_W;
  static constexpr int PLANE_B = PLANE_W * 4;
  static constexpr int SLAB_B = 2 * IW * 64;
  static constexpr int RING_B = (RING * SLAB_B + 1023) / 1024 * 1024;
  static constexpr int PLANES_B = CB * PLANE_B;
  static constexpr int WTAB_B = CB * WTAB_WORDS * 4;
  static constexpr int WSTAGE_B = (THREADS / 32) * 1024;
  static constexpr int BAR_B = 128;
  static constexpr int SMEM_B = RING_B + PLANES_B + WTAB_B + WSTAGE_B + BAR_B + 1024;
  static_assert((ROW_B / 16) % 2 == 1, "plane rows must be an odd number of 16-byte units");
  static_assert(SLAB_B % 512 == 0, "slabs must keep the 64-byte swizzle phase");
  static_assert(PLANE_W % 8 == 4, "plane pitch must be 4 (mod 8) words");
};

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, int c0, int c1, int c2, int c3,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], "
      "[%6];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, int c0, int c1, int c2, int c3, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void stsm_x4_trans(uint32_t addr, const uint32_t (&r)[4]) {
  asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3])
               : "memory");
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// physical plane row of tile-local input row y: even rows first, then odd rows, so that the eight rows
// {2r + j} an A-fragment tile needs are consecutive (conflict-free ldmatrix) for either parity of j
__device__ __forceinline__ int plane_row(int y) { return (y & 1) * (IH / 2) + (y >> 1); }

struct TileCoord { int c0, x0, y0, b; };

template <int NB>
__device__ __forceinline__ TileCoord decode_tile(int id, int n_cblk, int tiles_x, int tiles_y) {
  TileCoord tc;
  tc.c0 = (id % n_cblk) * CB;  // channel blocks of one spatial tile are neighbours in time: they share L2 lines
  id /= n_cblk;
  tc.x0 = (id % tiles_x) * Geo<NB>::TW;
  id /= tiles_x;
  tc.y0 = (id % tiles_y) * TH;
  tc.b = id / tiles_y;
  return tc;
}

// Persistent: a CTA walks tiles id = blockIdx.x, += gridDim.x.  The TMA loads of the NEXT tile's first RING
// slabs are issued as soon as this tile's slabs have been transposed into the planes, so they are in flight
// while the tensor-core and store phases of this tile run (ncu of the one-tile-per-CTA version: 25 % of all
// warp samples sat in the slab-arrival wait, ~20 KB in flight per SM against the ~65 KB HBM needs).
template <int NB>
__global__ void __launch_bounds__(THREADS, 2)
dwconv7_mma_kernel(const __grid_constant__ CUtensorMap tmap_in, const uint32_t* __restrict__ wtab,
                   const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int H, int W, int C,
                   int tiles_x, int tiles_y, int n_cblk, int total_tiles) {
  using G = Geo<NB>;
  extern __shared__ uint8_t smem_dwm[];
  const uint32_t base = (ptx::smem_u32(smem_dwm) + 1023u) & ~1023u;
  const uint32_t ring = base;
  const uint32_t planes = ring + G::RING_B;
  const uint32_t s_wtab = planes + G::PLANES_B;
  const uint32_t wstage = s_wtab + G::WTAB_B;   // per-warp 2 x 512 B transposition buffers of the writer
  const uint32_t bars = wstage + G::WSTAGE_B;
  auto full_bar = [&](int s) { return bars + 8u * s; };
  const uint32_t wbar = bars + 8u * NSLAB;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_in);
    for (int s = 0; s <= NSLAB; ++s) ptx::mbar_init(bars + 8u * s, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  pdl_sync();
  if (tid == 0 && static_cast<int>(blockIdx.x) < total_tiles) {
    const TileCoord tc = decode_tile<NB>(blockIdx.x, n_cblk, tiles_x, tiles_y);
    ptx::mbar_arrive_expect_tx(wbar, G::WTAB_B);
    bulk_g2s(s_wtab, wtab + static_cast<size_t>(tc.c0) * WTAB_WORDS, G::WTAB_B, wbar);
#pragma unroll
    for (int s = 0; s < RING; ++s) {
      ptx::mbar_arrive_expect_tx(full_bar(s), G::SLAB_B);
      tma_load_4d(ring + s * G::SLAB_B, &tmap_in, tc.c0, tc.x0 - 3, tc.y0 - 3 + 2 * s, tc.b, full_bar(s));
    }
  }

  // Toeplitz B fragment of kernel row ky: B[k][n] = w[ky][k - n]; this lane holds k = 2t(+1), 2t+8(+9), n = g.
  // Table word i of a kernel row = (w[i-1], w[i]) with w[-1] = w[7] = 0, word 8 = 0.
  const int i0 = 2 * t - g + 1, i1 = i0 + 8;
  const uint32_t o0 = static_cast<uint32_t>((i0 >= 0 && i0 <= 7) ? i0 : 8) * 4u;
  const uint32_t o1 = static_cast<uint32_t>((i1 >= 0 && i1 <= 7) ? i1 : 8) * 4u;
  // ldmatrix row address of this lane: tile j = lane >> 3 (kernel-row offset), row r = lane & 7
  const int lj = lane >> 3, lr = lane & 7;
  const uint32_t a_off = static_cast<uint32_t>(((lj & 1) * (IH / 2) + (lj >> 1) + lr) * G::ROW_B);

  uint32_t parity = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, parity ^= 1u) {
    const TileCoord tc = decode_tile<NB>(tile, n_cblk, tiles_x, tiles_y);
    const int next = tile + static_cast<int>(gridDim.x);

    // ---- NHWC slabs -> per-channel planes ----
    for (int s = warp; s < NSLAB; s += THREADS / 32) {
      const int buf = s % RING;
      ptx::mbar_wait(full_bar(s), parity);
      const uint32_t slab = ring + buf * G::SLAB_B;
#pragma unroll
      for (int task = 0; task < 2 * G::XB; ++task) {
        const int r = task / G::XB, xb = task % G::XB;
        const int px = r * G::IW + xb * 8 + (lane & 7);
        const int cv = lane >> 3;
        uint32_t R[4];
        ldsm_x4_trans(R, slab + px * 64 + ((cv ^ ((px >> 1) & 3)) << 4));
        const int y = 2 * s + r;
        const uint32_t dst = planes + g * G::PLANE_B + plane_row(y) * G::ROW_B + (xb * 4 + t) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) sts32(dst + q * 8 * G::PLANE_B, R[q]);
      }
      if (s + RING < NSLAB) {
        __syncwarp();
        if (lane == 0) {
          ptx::fence_proxy_async_smem();
          ptx::mbar_arrive_expect_tx(full_bar(s + RING), G::SLAB_B);
          tma_load_4d(slab, &tmap_in, tc.c0, tc.x0 - 3, tc.y0 - 3 + 2 * (s + RING), tc.b, full_bar(s + RING));
        }
      }
    }
    ptx::mbar_wait(wbar, parity);
    __syncthreads();  // planes complete, every slab of this tile consumed: the ring is free
    if (tid == 0 && next < total_tiles) {
      const TileCoord nt = decode_tile<NB>(next, n_cblk, tiles_x, tiles_y);
      ptx::fence_proxy_async_smem();
#pragma unroll
      for (int s = 0; s < RING; ++s) {
        ptx::mbar_arrive_expect_tx(full_bar(s), G::SLAB_B);
        tma_load_4d(ring + s * G::SLAB_B, &tmap_in, nt.c0, nt.x0 - 3, nt.y0 - 3 + 2 * s, nt.b, full_bar(s));
      }
    }

    // ---- tensor-core phase: each warp takes 4 channels ----
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const int ch = warp * 4 + i;
      const uint32_t plane = planes + ch * G::PLANE_B;
      const uint32_t wrow = s_wtab + ch * (WTAB_WORDS * 4);
      uint32_t b0[7], b1[7];
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        b0[ky] = lds32(wrow + ky * 36 + o0);
        b1[ky] = lds32(wrow + ky * 36 + o1);
      }
      const float bv = __ldg(bias + tc.c0 + ch);
      // A tiles T[j][xb]: rows {2r + j}, pixels 8xb .. 8xb+7
      uint32_t T[8][NB + 1];
#pragma unroll
      for (int xb = 0; xb <= NB; ++xb) {
        uint32_t lo[4], hi[4];
        ldsm_x4(lo, plane + a_off + xb * 16);
        ldsm_x4(hi, plane + a_off + 2 * G::ROW_B + xb * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) { T[j][xb] = lo[j]; T[4 + j][xb] = hi[j]; }
      }
      float acc[NB][4];
#pragma unroll
      for (int nb = 0; nb < NB; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[nb][e] = bv;
#pragma unroll
      for (int ky = 0; ky < 7; ++ky)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb)
          mma_bf16_16816(acc[nb], T[ky][nb], T[ky + 1][nb], T[ky][nb + 1], T[ky + 1][nb + 1], b0[ky], b1[ky]);
      // this warp is the only reader of plane `ch`: once its tiles are in registers the plane can take the
      // outputs (rows 2g / 2g+1 of the mma tile, 18-word rows: conflict-free for this access pattern)
      __syncwarp();
      const uint32_t op = plane + (2 * g) * (OROW_W * 4) + t * 4;
#pragma unroll
      for (int nb = 0; nb < NB; ++nb) {
        sts32(op + nb * 16, pack2(acc[nb][0], acc[nb][1]));
        sts32(op + OROW_W * 4 + nb * 16, pack2(acc[nb][2], acc[nb][3]));
      }
    }
    __syncthreads();  // output planes complete; the Toeplitz table is no longer read
    if (tid == 0 && next < total_tiles) {
      const TileCoord nt = decode_tile<NB>(next, n_cblk, tiles_x, tiles_y);
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive_expect_tx(wbar, G::WTAB_B);
      bulk_g2s(s_wtab, wtab + static_cast<size_t>(nt.c0) * WTAB_WORDS, G::WTAB_B, wbar);
    }

    // ---- output planes -> NHWC: 8 px x 32 ch blocks through a per-warp 512-byte buffer, 64-byte runs to HBM ----
    {
      const uint32_t mine = wstage + static_cast<uint32_t>(warp) * 1024u;
      const int spx = lane >> 2, schunk = lane & 3;  // this lane's 16 bytes of the transposed block
      int k = 0;
      for (int task = warp; task < TH * NB; task += THREADS / 32, ++k) {
        const int y = task / NB, xb = task % NB;
        uint32_t R[4];
        const uint32_t src = planes + g * G::PLANE_B + y * (OROW_W * 4) + (xb * 4 + t) * 4;
#pragma unroll
        for (int q = 0; q < 4; ++q) R[q] = lds32(src + q * 8 * G::PLANE_B);
        const uint32_t buf = mine + static_cast<uint32_t>(k & 1) * 512u;
        const int px = lane & 7, cv = lane >> 3;
        stsm_x4_trans(buf + px * 64 + ((cv ^ ((px >> 1) & 3)) << 4), R);
        __syncwarp();
        uint4 v;
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                     : "r"(buf + spx * 64 + ((schunk ^ ((spx >> 1) & 3)) << 4)));
        __nv_bfloat16* dst = out + ((static_cast<size_t>(tc.b) * H + tc.y0 + y) * W + tc.x0 + xb * 8 + spx) * C +
                             tc.c0 + schunk * 8;
        *reinterpret_cast<uint4*>(dst) = v;
      }
    }
    __syncthreads();  // the planes are rewritten by the next tile's transposition
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Warp-specialised pipeline, four output rows per fragment row ("R4").
//
// ncu of the kernel above (profiles/): L1/shared 73 % busy, tensor pipe 29 %, issue 38 %, and a fifth of all warp
// samples waiting at the block barriers between its three phases — with two CTAs per SM the transposition (LSU),
// tensor-core and writer phases mostly run one after the other: per tile ~2700 shared-memory cycles + ~1900 tensor
// cycles + ~3000 issue cycles add up to the ~6800 cycles a tile takes.  Here the three phases are three groups of
// warps of ONE persistent CTA per SM, working on three different tiles at any time (three plane buffers):
//
//   warp 3          TMA issuer: 19 two-row slabs per tile (own buffers, refilled as soon as a slab is transposed)
//   warps 0-2       slabs -> planes[b]         (ldmatrix.x4.trans + 4 STS.32 per 2 rows x 8 px x 16 ch)
//   warps 4-11      planes[b] -> 56 mma.sync per channel -> output planes (in place)
//   warps 12-15     output planes -> HBM       (one transposing x4 load -> 16 contiguous NHWC bytes per lane)
//
// linked by mbarriers only (no block barrier after the prologue).  Two further changes cut the shared-memory bytes per
// output: (i) fragment row g owns FOUR consecutive output rows (tile height 32): A tile T[j] holds image rows
// {4g + j}, j = 0..9, and serves both output row pairs, so a channel loads 10 x 5 tiles per 1024 outputs (3.2 bytes
// per output byte instead of 5.1), the Toeplitz fragments are fetched once per 1024 outputs and the halo shrinks from
// 22/16 to 38/32 rows; planes keep rows of equal (y mod 4) together so the eight rows of a tile stay consecutive.
// (ii) The writer no longer bounces through a staging buffer: ldmatrix.trans may take each of its eight rows from a
// DIFFERENT plane, so one x4 load picks, for eight pixels, the rows of channels {8q + 2i, 8q + 2i + 1} and every lane
// ends up with the sixteen contiguous NHWC bytes (eight channels) of one pixel.
// 16 channels per tile (32-byte pixels, SWIZZLE_32B slabs) keep a plane buffer at 48 KB.
namespace r4 {
constexpr int CB = 16;
constexpr int TH = 32, IH = TH + 6;
constexpr int SROWS = 8, NSLAB = (IH + SROWS - 1) / SROWS, NPAIR = IH / 2;   // TMA slabs of 8 rows (the last: 6 + 2 unused)
constexpr int NB = 4, TW = 8 * NB, IW = TW + 8, XB = IW / 8;
constexpr int ROW_B = IW * 2;                                   // 80 B: an odd number of 16-byte units
constexpr int PLANE_W = IH * ROW_B / 4 + 4;                     // 764 words = 4 (mod 8)
constexpr int PLANE_B = PLANE_W * 4;
constexpr int OCTET_SKEW = 64;                                  // bytes added to the planes of channels 8..15 (writer banks)
constexpr int PLANES_B = (CB * PLANE_B + OCTET_SKEW + 127) / 128 * 128;
constexpr int NBUF = 3;
constexpr int SLAB_B = SROWS * IW * CB * 2;                     // 8 rows x 40 px x 32 B
constexpr int SLABS_B = (NSLAB * SLAB_B + 1023) / 1024 * 1024;  // every slab of a tile has its own buffer
constexpr int WTAB_B = CB * WTAB_WORDS * 4;
constexpr int BAR_B = 512;
constexpr int P_WARPS = 3, W_WARPS = 4, M_WARPS = 16;            // + 1 TMA issuer; warpgroups: {P, P, P, issuer} {W x 4} {M x 16}
constexpr int LIGHT_WARPS = P_WARPS + 1 + W_WARPS;
constexpr int WS_THREADS = (LIGHT_WARPS + M_WARPS) * 32;
constexpr int STAGE_ROWS = 2, STAGE_B = STAGE_ROWS * TW * CB * 2;  // writer staging slice: 2 rows x 32 px x 32 B of NHWC output, two per warp
constexpr int SMEM_B = SLABS_B + NBUF * (PLANES_B + WTAB_B) + 2 * W_WARPS * STAGE_B + BAR_B + 1024;
constexpr int LIGHT_REGS = 48, M_REGS = 96;                        // setmaxnreg: 768 threads launch with 80 registers each
static_assert(LIGHT_WARPS % 4 == 0 && M_WARPS % 4 == 0, "setmaxnreg works on aligned groups of four warps");
static_assert(LIGHT_WARPS * (80 - LIGHT_REGS) >= M_WARPS * (M_REGS - 80), "register pool");
static_assert(SMEM_B <= 232448, "exceeds the 227 KB dynamic shared memory limit");
static_assert(TH == 4 * W_WARPS * STAGE_ROWS, "each writer warp stores four 2-row slices of a tile");
static_assert((ROW_B / 16) % 2 == 1 && PLANE_W % 8 == 4 && SLAB_B % 256 == 0, "bank layout");
static_assert(M_WARPS == CB && 8 * (2 * NSLAB + 4 * NBUF) <= BAR_B, "roles");

// physical plane row of tile-local input row y: rows of equal (y mod 4) are consecutive (class sizes 10, 10, 9, 9)
__device__ __forceinline__ int plane_row(int y) {
  const int q = y & 3;
  return q * 10 - (q == 3 ? 1 : 0) + (y >> 2);
}
__device__ __forceinline__ uint32_t plane_base(uint32_t planes, int ch, int skew = OCTET_SKEW) {
  return planes + static_cast<uint32_t>(ch * PLANE_B + (ch >> 3) * skew);
}

struct TileCoord { int c0, x0, y0, b; };
__device__ __forceinline__ TileCoord decode_tile(int id, int n_cblk, int tiles_x, int tiles_y, int tw = TW, int th = TH) {
  TileCoord tc;   // x0 / y0 in OUTPUT pixels
  tc.c0 = (id % n_cblk) * CB;
  id /= n_cblk;
  tc.x0 = (id % tiles_x) * tw;
  id /= tiles_x;
  tc.y0 = (id % tiles_y) * th;
  tc.b = id / tiles_y;
  return tc;
}

// ---- stride 2, two output channels per input channel (FastViTHD PatchEmbed: ReparamLargeKernelConv 7x7 s2, groups = C,
// 2C outputs, + GELU) on the same pipeline: the 38 x 40 x 16-channel input tile, its slabs and planes are those of the
// stride-1 kernel; a tile produces 16 x 16 output pixels x 32 output channels (out channel = 2 * c + m).
//   rows:    fragment row g owns output rows 2g, 2g+1 = input rows 4g + ky and 4g + 2 + ky: one x4 load of the A tiles
//            (ky, xb), (ky+2, xb), (ky, xb+1), (ky+2, xb+1) is the whole A operand of kernel row ky;
//   columns: output pixel n of an 8-pixel block reads input pixels 2n + kx: B[k][n] = w[ky][k - 2n], 21 inputs per block =
//            one k16 step + one k8 step (whose A operand is the first register pair of the next block's quad);
//   per channel 7 kernel rows x 2 blocks x 2 output channels x (k16 + k8) = 56 mma, 17 loads.
constexpr int S2_TWO = 16, S2_THO = 16;                 // output tile
constexpr int S2_OCB = 2 * CB;                          // output channels per tile
constexpr int S2_SKEW = 32;                             // plane skew of input channels 8..15 (writer banks, see below)
constexpr int S2_MOFF = S2_THO * ROW_B + 16;            // second output plane of an input channel, = 16 (mod 128) bytes
constexpr int S2_WZERO = 56;                            // table word that is always zero
constexpr int S2_STAGE_B = 2 * S2_TWO * S2_OCB * 2;     // writer staging slice: 2 rows x 16 px x 64 B
static_assert(2 * S2_STAGE_B <= STAGE_B * 2 && S2_MOFF + S2_THO * ROW_B <= PLANE_B, "stride-2 buffers fit the stride-1 ones");

__device__ __forceinline__ void mma_bf16_1688(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5}, {%6}, {%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(b0));
}

__device__ __forceinline__ void dw7_tensor_role_s2(uint32_t planes0, uint32_t wtab0, uint32_t bars,
                                                   const float* __restrict__ bias, int n_cblk, int total_tiles,
                                                   int act) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int lj = lane >> 3, lr = lane & 7;
  auto planes_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + b); };
  auto out_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + NBUF + b); };
  auto wtab_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + 3 * NBUF + b); };
  const int ch = warp - LIGHT_WARPS;
  // Toeplitz fragments: table word p of (m, ky) = taps (2p, 2p+1); this lane's pairs are p = t - g (k-lo), + 4 (k-hi), + 8 (k8 step)
  uint32_t o[3];
#pragma unroll
  for (int q = 0; q < 3; ++q) {
    const int pq = t - g + 4 * q;
    o[q] = static_cast<uint32_t>((pq >= 0 && pq <= 3) ? pq : -1);
  }
  // A quads of kernel row ky: lane address of matrix (lane >> 3): tile row class ky + 2 * (lj & 1), pixel block + (lj >> 1)
  uint32_t a_off[7];
#pragma unroll
  for (int ky = 0; ky < 7; ++ky)
    a_off[ky] = static_cast<uint32_t>((plane_row(ky + 2 * (lj & 1)) + lr) * ROW_B + (lj >> 1) * 16);
  uint32_t n = 0;
  for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n) {
    const int c0 = (tile % n_cblk) * CB;
    const uint32_t b = n % NBUF, k = n / NBUF;
    const float bv0 = __ldg(bias + 2 * (c0 + ch)), bv1 = __ldg(bias + 2 * (c0 + ch) + 1);
    const uint32_t plane = plane_base(planes0 + b * PLANES_B, ch, S2_SKEW);
    const uint32_t wrow = wtab0 + b * WTAB_B + ch * (WTAB_WORDS * 4);
    ptx::mbar_wait(wtab_full(b), k & 1u);
    ptx::mbar_wait(planes_full(b), k & 1u);
    float acc[2][2][4];   // [m][pixel block][fragment]
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[m][nb][e] = m ? bv1 : bv0;
#pragma unroll
    for (int ky = 0; ky < 7; ++ky) {
      uint32_t bw[2][3];
#pragma unroll
      for (int m = 0; m < 2; ++m)
#pragma unroll
        for (int q = 0; q < 3; ++q)
          bw[m][q] = lds32(wrow + ((o[q] == 0xffffffffu ? S2_WZERO : (m * 7 + ky) * 4 + static_cast<int>(o[q])) << 2));
      uint32_t A0[4], A1[4], A2[2];
      ldsm_x4(A0, plane + a_off[ky]);               // pixel blocks 0, 1
      ldsm_x4(A1, plane + a_off[ky] + 32);          // pixel blocks 2, 3
      asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0, %1}, [%2];"
                   : "=r"(A2[0]), "=r"(A2[1])
                   : "r"(plane + a_off[ky] - (lj >> 1) * 16 + 64));   // pixel block 4 (lanes 0-15 address it)
#pragma unroll
      for (int m = 0; m < 2; ++m) {
        mma_bf16_16816(acc[m][0], A0[0], A0[1], A0[2], A0[3], bw[m][0], bw[m][1]);
        mma_bf16_1688(acc[m][0], A1[0], A1[1], bw[m][2]);
        mma_bf16_16816(acc[m][1], A1[0], A1[1], A1[2], A1[3], bw[m][0], bw[m][1]);
        mma_bf16_1688(acc[m][1], A2[0], A2[1], bw[m][2]);
      }
    }
    // outputs (+ GELU) -> the channel's own plane: two 16 x 16 output planes, 80-byte rows, chunk ^ (row >> 3)
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int nb = 0; nb < 2; ++nb) {
        float v[4] = {acc[m][nb][0], acc[m][nb][1], acc[m][nb][2], acc[m][nb][3]};
        if (act == ACT_GELU) {
#pragma unroll
          for (int e = 0; e < 4; ++e) v[e] = gelu_tanh_fit(v[e]);
        }
        const uint32_t dst = plane + m * S2_MOFF + (2 * g) * ROW_B + ((nb ^ (g >> 2)) << 4) + t * 4;
        sts32(dst, pack2(v[0], v[1]));
        sts32(dst + ROW_B, pack2(v[2], v[3]));
      }
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(out_full(b));
  }
}

// tensor-core role of the kernel below (one channel per warp); a function of its own so that it is compiled against
// the register budget its setmaxnreg.inc grants
__device__ __forceinline__ void dw7_tensor_role(uint32_t planes0, uint32_t wtab0, uint32_t bars,
                                                const float* __restrict__ bias, int n_cblk, int total_tiles, int debug) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int lj = lane >> 3, lr = lane & 7;
  auto planes_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + b); };
  auto out_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + NBUF + b); };
  auto wtab_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + 3 * NBUF + b); };
  {
    // ===================== tensor cores: one channel per warp =====================
    const int ch = warp - LIGHT_WARPS;
    // Toeplitz B fragment of kernel row ky (same table as the kernel above)
    const int i0 = 2 * t - g + 1, i1 = i0 + 8;
    const uint32_t o0 = static_cast<uint32_t>((i0 >= 0 && i0 <= 7) ? i0 : 8) * 4u;
    const uint32_t o1 = static_cast<uint32_t>((i1 >= 0 && i1 <= 7) ? i1 : 8) * 4u;
    // One x4 load = one A operand, no register shuffling: matrices (J, nb), (J+1, nb), (J, nb+1), (J+1, nb+1) for EVEN J.
    // Fragment row g owns output rows 4g .. 4g+3.  The quad of tile rows (J, J+1) times kernel row ky lands on the output
    // row pair starting at J - ky, so even kernel rows feed the pairs (0,1) and (2,3) and odd ones the pairs (1,2),
    // (-1,0) and (3,4), whose out-of-range halves are dropped: 17 instead of 14 mma per 8-pixel block, but each is fed by
    // one of 5 loads instead of four register moves (SASS of the shuffling version: 220 moves for 56 mma).
    uint32_t a_off[5];
#pragma unroll
    for (int q = 0; q < 5; ++q)
      a_off[q] = static_cast<uint32_t>((plane_row(2 * q + (lj & 1)) + lr) * ROW_B + (lj >> 1) * 16);
    const uint32_t o_off = static_cast<uint32_t>(4 * g * ROW_B + t * 4);
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n) {
      const int c0 = (tile % n_cblk) * CB;
      const uint32_t b = n % NBUF, k = n / NBUF;
      const float bv = __ldg(bias + c0 + ch);
      const uint32_t plane = plane_base(planes0 + b * PLANES_B, ch);
      const uint32_t wrow = wtab0 + b * WTAB_B + ch * (WTAB_WORDS * 4);
      ptx::mbar_wait(wtab_full(b), k & 1u);
      uint32_t b0[7], b1[7];
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
        b0[ky] = lds32(wrow + ky * 36 + o0);
        b1[ky] = lds32(wrow + ky * 36 + o1);
      }
      ptx::mbar_wait(planes_full(b), k & 1u);
      if (debug & 1) { __syncwarp(); if (lane == 0) ptx::mbar_arrive(out_full(b)); continue; }
      uint32_t packed[NB][4];   // [pixel block][output row 4g + i]
#pragma unroll
      for (int h2 = 0; h2 < NB / 2; ++h2) {
        float E0[2][4], E2[2][4], O1[2][4], Om[2][4], O3[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int e = 0; e < 4; ++e) { E0[u][e] = bv; E2[u][e] = bv; O1[u][e] = 0.f; Om[u][e] = 0.f; O3[u][e] = 0.f; }
#pragma unroll
        for (int q = 0; q < 5; ++q) {
          const int J = 2 * q;
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            uint32_t A[4];
            ldsm_x4(A, plane + a_off[q] + (2 * h2 + u) * 16);
            if (J <= 6) mma_bf16_16816(E0[u], A[0], A[1], A[2], A[3], b0[J > 6 ? 0 : J], b1[J > 6 ? 0 : J]);
            if (J >= 2) mma_bf16_16816(E2[u], A[0], A[1], A[2], A[3], b0[J >= 2 ? J - 2 : 0], b1[J >= 2 ? J - 2 : 0]);
            if (J >= 2 && J <= 6) mma_bf16_16816(O1[u], A[0], A[1], A[2], A[3], b0[J >= 2 && J <= 6 ? J - 1 : 0], b1[J >= 2 && J <= 6 ? J - 1 : 0]);
            if (J <= 4) mma_bf16_16816(Om[u], A[0], A[1], A[2], A[3], b0[J <= 4 ? J + 1 : 0], b1[J <= 4 ? J + 1 : 0]);
            if (J >= 4) mma_bf16_16816(O3[u], A[0], A[1], A[2], A[3], b0[J >= 4 ? J - 3 : 0], b1[J >= 4 ? J - 3 : 0]);
          }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          packed[2 * h2 + u][0] = pack2(E0[u][0] + Om[u][2], E0[u][1] + Om[u][3]);
          packed[2 * h2 + u][1] = pack2(E0[u][2] + O1[u][0], E0[u][3] + O1[u][1]);
          packed[2 * h2 + u][2] = pack2(E2[u][0] + O1[u][2], E2[u][1] + O1[u][3]);
          packed[2 * h2 + u][3] = pack2(E2[u][2] + O3[u][0], E2[u][3] + O3[u][1]);
        }
      }
      // this warp is the only reader of plane `ch`: once its tiles are in registers the plane takes the outputs
      // (natural row order, 80-byte rows, 16-byte chunk index XOR (y >> 3): conflict-free for the fragment stores)
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t orow = plane + o_off + i * ROW_B;
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) sts32(orow + ((nb ^ (g >> 1)) << 4), packed[nb][i]);
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(out_full(b));
    }
  }
}

// S2 = false: stride 1 (debug = stage-skipping bits); S2 = true: stride 2 x 2 output channels (debug = activation)
template <bool S2>
__global__ void __launch_bounds__(WS_THREADS, 1)
dwconv7_mma_r4_kernel(const __grid_constant__ CUtensorMap tmap_in, const __grid_constant__ CUtensorMap tmap_out,
                      const uint32_t* __restrict__ wtab, const float* __restrict__ bias, int H, int W, int C,
                      int tiles_x, int tiles_y, int n_cblk, int total_tiles, int debug) {
  constexpr int TWx = S2 ? S2_TWO : TW, THx = S2 ? S2_THO : TH, SC = S2 ? 2 : 1;
  extern __shared__ uint8_t smem_dwm[];
  const uint32_t base = (ptx::smem_u32(smem_dwm) + 1023u) & ~1023u;
  const uint32_t slabs = base;
  const uint32_t stage0 = slabs + SLABS_B;
  const uint32_t planes0 = stage0 + 2 * W_WARPS * STAGE_B;
  const uint32_t wtab0 = planes0 + NBUF * PLANES_B;
  const uint32_t bars = wtab0 + NBUF * WTAB_B;
  auto slab_full = [&](int s) { return bars + 8u * s; };
  auto slab_empty = [&](int s) { return bars + 8u * (NSLAB + s); };
  auto planes_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + b); };
  auto out_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + NBUF + b); };
  auto planes_empty = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + 2 * NBUF + b); };
  auto wtab_full = [&](uint32_t b) { return bars + 8u * (2 * NSLAB + 3 * NBUF + b); };

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int lj = lane >> 3, lr = lane & 7;

  if (tid == 0) {
    ptx::prefetch_tmap(&tmap_in);
    ptx::prefetch_tmap(&tmap_out);
    for (int s = 0; s < NSLAB; ++s) {
      ptx::mbar_init(slab_full(s), 1);
      ptx::mbar_init(slab_empty(s), (IH - SROWS * s < SROWS ? IH - SROWS * s : SROWS) / 2);   // one arrival per row pair
    }
    for (uint32_t b = 0; b < NBUF; ++b) {
      ptx::mbar_init(planes_full(b), P_WARPS);
      ptx::mbar_init(out_full(b), M_WARPS);
      ptx::mbar_init(planes_empty(b), W_WARPS);
      ptx::mbar_init(wtab_full(b), 1);
    }
    ptx::fence_barrier_init();
  }
  __syncthreads();
  pdl_sync();

  if (warp >= LIGHT_WARPS) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(M_REGS));
    if constexpr (S2) dw7_tensor_role_s2(planes0, wtab0, bars, bias, n_cblk, total_tiles, debug);
    else dw7_tensor_role(planes0, wtab0, bars, bias, n_cblk, total_tiles, debug);
    return;
  }
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(LIGHT_REGS));
  if (warp == P_WARPS) {
    // ===================== TMA issuer =====================
    if (lane == 0) {
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n) {
        const TileCoord tc = decode_tile(tile, n_cblk, tiles_x, tiles_y, TWx, THx);
        const uint32_t b = n % NBUF, k = n / NBUF;
        ptx::mbar_wait(out_full(b), (k & 1u) ^ 1u);       // the tensor-core warps of tile n-3 are done with wtab[b]
        ptx::mbar_arrive_expect_tx(wtab_full(b), WTAB_B);
        bulk_g2s(wtab0 + b * WTAB_B, wtab + static_cast<size_t>(tc.c0) * WTAB_WORDS, WTAB_B, wtab_full(b));
#pragma unroll 1
        for (int s = 0; s < NSLAB; ++s) {
          ptx::mbar_wait(slab_empty(s), (n & 1u) ^ 1u);
          ptx::mbar_arrive_expect_tx(slab_full(s), SLAB_B);
          tma_load_4d(slabs + s * SLAB_B, &tmap_in, tc.c0, SC * tc.x0 - 3, SC * tc.y0 - 3 + SROWS * s, tc.b, slab_full(s));
        }
      }
    }
  } else if (warp < P_WARPS) {
    // ===================== NHWC slabs -> per-channel planes =====================
    // x4 matrix i = lane >> 3: slab row i >> 1, channel octet i & 1; this lane addresses pixel (lane & 7) of the block
    const int mrow = lj >> 1, cv = lj & 1;
    uint32_t src_off[XB], dst_off[XB];
#pragma unroll
    for (int xb = 0; xb < XB; ++xb) {
      const int px = mrow * IW + xb * 8 + lr;
      src_off[xb] = static_cast<uint32_t>(px * 32 + ((cv ^ ((px >> 2) & 1)) << 4));
      dst_off[xb] = static_cast<uint32_t>((xb * 4 + t) * 4);
    }
    const uint32_t ch_off0 = static_cast<uint32_t>(g * PLANE_B);
    const uint32_t ch_off1 = static_cast<uint32_t>((8 + g) * PLANE_B + (S2 ? S2_SKEW : OCTET_SKEW));
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n) {
      const uint32_t b = n % NBUF, k = n / NBUF;
      const uint32_t planes = planes0 + b * PLANES_B;
      ptx::mbar_wait(planes_empty(b), (k & 1u) ^ 1u);
#pragma unroll 1
      for (int rp = warp; rp < NPAIR; rp += P_WARPS) {   // task = one row pair of a slab
        const int s = rp / (SROWS / 2);
        ptx::mbar_wait(slab_full(s), n & 1u);
        const uint32_t slab = slabs + s * SLAB_B + (rp % (SROWS / 2)) * (2 * IW * CB * 2);
        const uint32_t r0 = planes + static_cast<uint32_t>(plane_row(2 * rp) * ROW_B);
        const uint32_t r1 = planes + static_cast<uint32_t>(plane_row(2 * rp + 1) * ROW_B);
        if (debug & 2) { __syncwarp(); if (lane == 0) ptx::mbar_arrive(slab_empty(s)); continue; }
        uint32_t R[XB][4];   // all five loads in flight before the first store (one shared-memory latency per task)
#pragma unroll
        for (int xb = 0; xb < XB; ++xb) ldsm_x4_trans(R[xb], slab + src_off[xb]);
#pragma unroll
        for (int xb = 0; xb < XB; ++xb) {
          sts32(r0 + ch_off0 + dst_off[xb], R[xb][0]);
          sts32(r0 + ch_off1 + dst_off[xb], R[xb][1]);
          sts32(r1 + ch_off0 + dst_off[xb], R[xb][2]);
          sts32(r1 + ch_off1 + dst_off[xb], R[xb][3]);
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(slab_empty(s));
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(planes_full(b));
    }
  } else {
    if constexpr (S2) {
      // ===================== writer, stride-2 kernel: 32 output channels, 64 bytes per pixel =====================
      // matrix i = lane >> 3 holds the channel pair 2i, 2i+1 of each of the four octets: its row r is output channel
      // co = 8 * (r >> 1) + 2i + (r & 1), i.e. plane (co >> 1), half (co & 1); lane (g, t) receives octet t of pixel g.
      // Banks: planes step by 7 (mod 8) 16-byte units, the octet skew adds 2 per 8 input channels, the second output
      // plane 1: the eight rows of a matrix land in eight different bank groups.
      const int ww = warp - (P_WARPS + 1);
      const int co = 8 * (lr >> 1) + 2 * lj + (lr & 1);
      const uint32_t w_rel = static_cast<uint32_t>((co >> 1) * PLANE_B + ((co >> 4) & 1) * S2_SKEW + (co & 1) * S2_MOFF);
      const uint32_t stage_w = stage0 + ww * 2 * STAGE_B;
      const uint32_t st_off = static_cast<uint32_t>(g * 64 + t * 16);
      uint32_t n = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n) {
        const TileCoord tc = decode_tile(tile, n_cblk, tiles_x, tiles_y, TWx, THx);
        const uint32_t b = n % NBUF, k = n / NBUF;
        const uint32_t w_plane = planes0 + b * PLANES_B + w_rel;
        ptx::mbar_wait(out_full(b), k & 1u);
#pragma unroll 1
        for (int r = 0; r < 2; ++r) {
          const int ybase = 4 * ww + 2 * r;
          const uint32_t stage = stage_w + (r & 1) * S2_STAGE_B;
          uint32_t R[4][4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int y = ybase + (q >> 1), xb = q & 1;
            ldsm_x4_trans(R[q], w_plane + y * ROW_B + ((xb ^ ((y >> 3) & 1)) << 4));
          }
          if (lane == 0) ptx::tma_store_wait_read<1>();
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + st_off + ((q >> 1) * S2_TWO + (q & 1) * 8) * 64),
                         "r"(R[q][0]), "r"(R[q][1]), "r"(R[q][2]), "r"(R[q][3])
                         : "memory");
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_4d(&tmap_out, 2 * tc.c0, tc.x0, tc.y0 + ybase, tc.b, stage);
            ptx::tma_store_commit();
          }
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(planes_empty(b));
      }
      if (lane == 0) ptx::tma_store_wait<0>();
    } else {
    // ===================== writer: output planes -> NHWC =====================
    // matrix i = lane >> 3 holds the channel pair 2i, 2i+1 of each octet; its row r = lane & 7 is
    // (channel 8*((r >> 1) & 1) + 2i + (r & 1), pixel block +2*(r >> 2)); task = (row, pixel blocks {xb0, xb0+2}).
    // Each lane ends up with the 16 contiguous NHWC bytes (8 channels) of one pixel; they go through a staging
    // slice and out with one TMA store per two rows (ncu: scattered 32-byte runs straight from registers cost ~21 LSU wavefronts
    // per STG.128, 28 % of the kernel's LSU traffic).
    const int ww = warp - (P_WARPS + 1);
    const int w_ch = 8 * ((lr >> 1) & 1) + 2 * lj + (lr & 1), w_pb = 2 * (lr >> 2);
    const int pb = 2 * (t >> 1), oct = t & 1;
    const uint32_t stage_w = stage0 + ww * 2 * STAGE_B;
    const uint32_t st_off = static_cast<uint32_t>(((pb * 8 + g) * CB + oct * 8) * 2);
    uint32_t n = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n) {
      const TileCoord tc = decode_tile(tile, n_cblk, tiles_x, tiles_y);
      const uint32_t b = n % NBUF, k = n / NBUF;
      const uint32_t w_plane = plane_base(planes0 + b * PLANES_B, w_ch);
      ptx::mbar_wait(out_full(b), k & 1u);
      if (debug & 4) { __syncwarp(); if (lane == 0) ptx::mbar_arrive(planes_empty(b)); continue; }
#pragma unroll 1
      for (int r = 0; r < 4; ++r) {
        const int ybase = (4 * ww + r) * STAGE_ROWS;
        const uint32_t stage = stage_w + (r & 1) * STAGE_B;
        uint32_t R[2 * STAGE_ROWS][4];
#pragma unroll
        for (int q = 0; q < 2 * STAGE_ROWS; ++q) {
          const int y = ybase + (q >> 1), xb0 = q & 1;
          ldsm_x4_trans(R[q], w_plane + y * ROW_B + (((xb0 + w_pb) ^ ((y >> 3) & 3)) << 4));
        }
        if (lane == 0) ptx::tma_store_wait_read<1>();   // the slice stored two rounds ago has left this staging buffer
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 2 * STAGE_ROWS; ++q)   // dense box (TMA sources are 128-byte aligned, a swizzled map pads 32-byte rows): the two
                                                   // pixel blocks of a lane quad are 512 B apart and share banks, 2-way
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage + st_off + ((q >> 1) * TW + (q & 1) * 8) * CB * 2),
                       "r"(R[q][0]), "r"(R[q][1]), "r"(R[q][2]), "r"(R[q][3])
                       : "memory");
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_4d(&tmap_out, tc.c0, tc.x0, tc.y0 + ybase, tc.b, stage);
          ptx::tma_store_commit();
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(planes_empty(b));
    }
    if (lane == 0) ptx::tma_store_wait<0>();
    }
  }
}
}  // namespace r4

// word i (0..8) of kernel row ky for channel c: (w[ky][i-1], w[ky][i]) as bf16, zero outside 0..6
__global__ void dwconv7_wtab_kernel(const float* __restrict__ w, int C, uint32_t* __restrict__ wtab) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * WTAB_WORDS) return;
  const int c = idx / WTAB_WORDS, j = idx % WTAB_WORDS;
  uint32_t v = 0;
  if (j < 63) {
    const int ky = j / 9, i = j % 9;
    const float lo = (i >= 1 && i <= 7) ? w[static_cast<size_t>(ky * 7 + i - 1) * C + c] : 0.0f;
    const float hi = (i <= 6) ? w[static_cast<size_t>(ky * 7 + i) * C + c] : 0.0f;
    v = pack2(lo, hi);
  }
  wtab[idx] = v;
}

// stride-2 x2 table: word (m * 7 + ky) * 4 + p of input channel c = taps (2p, 2p+1) of kernel row ky of output channel
// 2c + m as bf16 (tap 7 = 0); words 56..63 = 0
__global__ void dwconv7_s2_wtab_kernel(const float* __restrict__ w, int C, uint32_t* __restrict__ wtab) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * WTAB_WORDS) return;
  const int c = idx / WTAB_WORDS, j = idx % WTAB_WORDS;
  uint32_t v = 0;
  if (j < 56) {
    const int p = j & 3, mk = j >> 2, m = mk / 7, ky = mk % 7;
    const size_t co = static_cast<size_t>(2 * c + m), ld = static_cast<size_t>(2 * C);
    const float lo = w[static_cast<size_t>(ky * 7 + 2 * p) * ld + co];
    const float hi = (2 * p + 1 <= 6) ? w[static_cast<size_t>(ky * 7 + 2 * p + 1) * ld + co] : 0.0f;
    v = pack2(lo, hi);
  }
  wtab[idx] = v;
}

int make_tmap_nhwc(CUtensorMap* out, const void* ptr, int B, int H, int W, int C, int box_w, int box_h,
                   int box_c = CB, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_64B) {
  TmaEncodeTiledFn fn = tma_encode_fn();
  FVLA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FVLA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA base must be 16-byte aligned");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(B)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2,
                           static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1u};
  cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (NHWC) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

template <int NB>
int launch_mma(const void* in, const uint32_t* wtab, const float* bias, void* out, int B, int H, int W, int C,
               cudaStream_t stream) {
  using G = Geo<NB>;
  auto kfn = dwconv7_mma_kernel<NB>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), G::SMEM_B)) return rc;
  CUtensorMap ti;
  if (int rc = make_tmap_nhwc(&ti, in, B, H, W, C, G::IW, 2)) return rc;
  const int tiles_x = W / G::TW, tiles_y = H / TH, n_cblk = C / CB;
  const long long total = static_cast<long long>(tiles_x) * tiles_y * n_cblk * B;
  FVLA_REQUIRE(total < (1ll << 31), "dwconv7_mma: too many tiles");
  const int resident = 2 * num_sms();  // two CTAs per SM (shared memory), each walking a stride of tiles
  const int grid = total < resident ? static_cast<int>(total) : resident;
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(THREADS), G::SMEM_B, stream, ti, wtab, bias,
                             static_cast<__nv_bfloat16*>(out), H, W, C, tiles_x, tiles_y, n_cblk,
                             static_cast<int>(total)));
  return 0;
}

int launch_mma_r4(const void* in, const uint32_t* wtab, const float* bias, void* out, int B, int H, int W, int C,
                  cudaStream_t stream) {
  auto kfn = r4::dwconv7_mma_r4_kernel<false>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), r4::SMEM_B)) return rc;
  CUtensorMap ti;
  if (int rc = make_tmap_nhwc(&ti, in, B, H, W, C, r4::IW, r4::SROWS, r4::CB, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
  CUtensorMap to;
  if (int rc = make_tmap_nhwc(&to, out, B, H, W, C, r4::TW, r4::STAGE_ROWS, r4::CB, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  const int tiles_x = W / r4::TW, tiles_y = H / r4::TH, n_cblk = C / r4::CB;
  const long long total = static_cast<long long>(tiles_x) * tiles_y * n_cblk * B;
  FVLA_REQUIRE(total < (1ll << 31), "dwconv7_mma: too many tiles");
  static const int dbg = std::getenv("FVLA_DW7_DEBUG") ? std::atoi(std::getenv("FVLA_DW7_DEBUG")) : 0;  // stage-skipping bits (timing only)
  const int resident = num_sms();   // one warp-specialised CTA per SM
  const int grid = total < resident ? static_cast<int>(total) : resident;
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(r4::WS_THREADS), r4::SMEM_B, stream, ti, to, wtab, bias, H, W, C,
                             tiles_x, tiles_y, n_cblk, static_cast<int>(total), dbg));
  return 0;
}

int launch_mma_s2(const void* in, const uint32_t* wtab, const float* bias, void* out, int B, int H, int W, int C, int act,
                  cudaStream_t stream) {
  auto kfn = r4::dwconv7_mma_r4_kernel<true>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), r4::SMEM_B)) return rc;
  CUtensorMap ti, to;
  if (int rc = make_tmap_nhwc(&ti, in, B, H, W, C, r4::IW, r4::SROWS, r4::CB, CU_TENSOR_MAP_SWIZZLE_32B)) return rc;
  if (int rc = make_tmap_nhwc(&to, out, B, H / 2, W / 2, 2 * C, r4::S2_TWO, 2, r4::S2_OCB, CU_TENSOR_MAP_SWIZZLE_NONE)) return rc;
  const int tiles_x = (W / 2) / r4::S2_TWO, tiles_y = (H / 2) / r4::S2_THO, n_cblk = C / r4::CB;
  const long long total = static_cast<long long>(tiles_x) * tiles_y * n_cblk * B;
  FVLA_REQUIRE(total < (1ll << 31), "dwconv7_s2m2_mma: too many tiles");
  const int resident = num_sms();
  const int grid = total < resident ? static_cast<int>(total) : resident;
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(r4::WS_THREADS), r4::SMEM_B, stream, ti, to, wtab, bias, H, W, C,
                             tiles_x, tiles_y, n_cblk, static_cast<int>(total), act));
  return 0;
}

}  // namespace

bool dwconv7_s2m2_mma_supported(int dtype, int H, int W, int C, int mult, int k, int stride, int act) {
  static const bool on = std::getenv("FVLA_DISABLE_DWCONV7_S2_MMA") == nullptr;  // A/B switch for profiling
  return on && dtype == DT_BF16 && k == 7 && stride == 2 && mult == 2 && (act == ACT_NONE || act == ACT_GELU) &&
         C % r4::CB == 0 && H % (2 * r4::S2_THO) == 0 && W % (2 * r4::S2_TWO) == 0;
}

int dwconv7_s2m2_mma_prepare(const float* w_packed, int C, uint32_t* wtab, cudaStream_t stream) {
  const int n = C * WTAB_WORDS;
  dwconv7_s2_wtab_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(w_packed, C, wtab);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int dwconv7_s2m2_mma(const void* in, const uint32_t* wtab, const float* bias, void* out, int B, int H, int W, int C,
                     int act, cudaStream_t stream) {
  return launch_mma_s2(in, wtab, bias, out, B, H, W, C, act, stream);
}

bool dwconv7_mma_supported(int dtype, int H, int W, int C, int mult, int k, int stride, int act) {
  return dtype == DT_BF16 && k == 7 && stride == 1 && mult == 1 && act == ACT_NONE && C % CB == 0 && H % TH == 0 &&
         (W % 32 == 0 || W == 16);
}

size_t dwconv7_wtab_bytes(int C) { return static_cast<size_t>(C) * WTAB_WORDS * 4; }

int dwconv7_mma_prepare(const float* w_packed, int C, uint32_t* wtab, cudaStream_t stream) {
  const int n = C * WTAB_WORDS;
  dwconv7_wtab_kernel<<<ceil_div(n, 256), 256, 0, stream>>>(w_packed, C, wtab);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int dwconv7_mma(const void* in, const uint32_t* wtab, const float* bias, void* out, int B, int H, int W, int C,
                cudaStream_t stream) {
  static const bool r4_on = std::getenv("FVLA_DISABLE_DWCONV7_R4") == nullptr;  // A/B switch for profiling
  if (r4_on && W % r4::TW == 0 && H % r4::TH == 0 && C % r4::CB == 0)
    return launch_mma_r4(in, wtab, bias, out, B, H, W, C, stream);
  if (W % 32 == 0) return launch_mma<4>(in, wtab, bias, out, B, H, W, C, stream);
  return launch_mma<2>(in, wtab, bias, out, B, H, W, C, stream);
}

}  // namespace fvla
