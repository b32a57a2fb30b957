// Fused ConvFFN tail for the C = 384 stage of FastViTHD:   out = resid + fc2( GELU( fc1(z) ) )
// (reference: HF-hub FastViTHD `convffn.fc1 -> act -> fc2`, layer scale folded into fc2).
//
// Unfused (two GEMMs through a 4x hidden tensor of 402 MB per 32 images) this is 22 % of the batch-64 step.  The
// stage-0/1 kernel (ffn_fused_sm100.cu) cannot be stretched to C = 384: its O accumulator would take 384 of the 512
// TMEM columns and leave no room for two 128-wide S buffers and H.  This kernel trades chunk width for columns:
//
//   TMEM (per CTA, 128 lanes):  O [0, 384)   S [384, 448)   H0 [448, 480)   H1 [480, 512)
//   per 256-row tile (CTA pair, cta_group::2; each CTA owns 128 rows):
//     X tile [128 x 384] bf16 resident in shared memory (six 64-wide K panels, one TMA load each)
//     for each 64-wide hidden chunk j:
//       GEMM1  S (128 x 64 fp32)  = X . W1[j]^T                  24 MMAs of N = 64; W1 chunks through a 2-stage TMA ring
//       epilogue warps: S -> registers (S is free again) -> +b1 -> GELU on packed half pairs -> H[j & 1] in TENSOR MEMORY
//                       (tcgen05.st; one 32-bit column per pair of hidden units)
//       GEMM2  O (128 x 384 fp32) += H[j & 1] . W2[:, j]^T       A from tensor memory, two N = 192 halves x 4 K steps;
//                                                                W2 chunks (fp16) through their own 2-stage ring
//     epilogue warps: O -> +b2 -> +resid -> bf16 -> global (64 contiguous bytes per lane and 32-column block)
//
// GEMM1 and GEMM2 are issued by two warps that meet only through the epilogue (S -> GELU -> H): GEMM1(j+1) starts as
// soon as S(j) is in registers — S is single-buffered, but that happens long before GEMM2(j) retires — and GELU(j) runs
// under GEMM1(j+1); H is double-buffered because GELU(j+1) finishes under GEMM2(j).
// Shared-memory traffic per chunk and CTA: 120 KB of GEMM1 operands + 24 KB of GEMM2 operands + 48 KB of landing
// weights against 1 536 tensor cycles; the hidden tensor never leaves the SM pair: algorithmic HBM traffic is
// read z + read resid + write out = 6 M C bytes.
#include "common.cuh"
#include "epilogue_math.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tma_host.h"

#include <cstdlib>

namespace fvla {
namespace {
using namespace epi;

constexpr int WM = 128;             // rows per CTA
constexpr int WPAIR_M = 256;        // rows per CTA pair
constexpr int WHC = 64;             // hidden columns per chunk
constexpr int W_THREADS = 640;      // warp 0: X / W1 producer, warp 1: GEMM1 issuer, warp 2: TMEM + W2 producer, warp 3: GEMM2 issuer, warps 4..19: epilogue
constexpr int W_EPI_WARPS = 16;
constexpr int WX_PANEL = WM * 128;  // 128 rows x 64 bf16, SWIZZLE_128B

template <int C> struct WideCfg {
  static constexpr int KP = C / 64;                          // K panels of X / W1
  static constexpr int X_BYTES = KP * WX_PANEL;
  static constexpr int W1_PANEL = (WHC / 2) * 128;           // this CTA's half of a W1 chunk panel: 32 rows x 128 B
  static constexpr int W1_STAGE = KP * W1_PANEL;             // one stage = the W1 rows of a whole chunk
  static constexpr int NH = C / 192;                         // GEMM2 runs as NH instructions of N = 192
  static constexpr int W2_HALF = (192 / 2) * 128;            // this CTA's 96 rows of one N = 192 half
  static constexpr int W2_STAGE = NH * W2_HALF;
  static constexpr int R1 = 2, R2 = 2;
  static constexpr int BAR_BYTES = 256;
  static constexpr int MAX_HIDDEN = 4 * C;
  static constexpr int BIAS_BYTES = C * 4 + MAX_HIDDEN * 2;  // b2 fp32, b1 as packed halves
  static constexpr int SMEM_BYTES = X_BYTES + R1 * W1_STAGE + R2 * W2_STAGE + BAR_BYTES + BIAS_BYTES + 1024;
  static constexpr uint32_t TMEM_S = C, TMEM_H = C + WHC;
  static_assert(C % 192 == 0 && C + WHC + 2 * (WHC / 2) <= 512, "O + S + two H buffers must fit the 512 TMEM columns");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

struct WideParams {
  int M, hidden;
  const float* b1;              // [hidden] (pre-halved with W1: the GELU epilogue takes x/2)
  const float* b2;              // [C]
  const __nv_bfloat16* resid;   // [M, C], may alias out
  __nv_bfloat16* out;           // [M, C]
  int dbg;                      // FVLA_FFN_WIDE_DEBUG bits (timing only, results garbage): 1 = no GEMM1 MMAs, 2 = no GEMM2 MMAs, 4 = no drain
};

template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(W_THREADS, 1)
ffn_wide_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                const __grid_constant__ CUtensorMap tmap_w2, const WideParams p) {
  using Cfg = WideCfg<C>;
  constexpr int R1 = Cfg::R1, R2 = Cfg::R2, KP = Cfg::KP, NH = Cfg::NH;
  extern __shared__ uint8_t smem_wide[];
  const uint32_t smem_base = (ptx::smem_u32(smem_wide) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;
  const uint32_t smem_w1 = smem_x + Cfg::X_BYTES;
  const uint32_t smem_w2 = smem_w1 + R1 * Cfg::W1_STAGE;
  const uint32_t bar_base = smem_w2 + R2 * Cfg::W2_STAGE;
  auto w1full = [&](int s) { return bar_base + 8u * s; };
  auto w1empty = [&](int s) { return bar_base + 8u * (R1 + s); };
  auto w2full = [&](int s) { return bar_base + 8u * (2 * R1 + s); };
  auto w2empty = [&](int s) { return bar_base + 8u * (2 * R1 + R2 + s); };
  const uint32_t xfull = bar_base + 8u * (2 * R1 + 2 * R2), xempty = xfull + 8u;
  const uint32_t sfull = xempty + 8u, sempty = xempty + 16u;
  auto hfull = [&](uint32_t b) { return xempty + 24u + 8u * b; };
  auto hempty = [&](uint32_t b) { return xempty + 40u + 8u * b; };
  const uint32_t ofull = xempty + 56u, oempty = xempty + 64u;
  const uint32_t tmem_ptr_smem = xempty + 72u;
  static_assert(8 * (2 * R1 + 2 * R2 + 1) + 72 + 4 <= Cfg::BAR_BYTES, "barrier block too small");
  float* s_b2 = reinterpret_cast<float*>(smem_wide + (bar_base - ptx::smem_u32(smem_wide)) + Cfg::BAR_BYTES);
  uint32_t* s_b1h = reinterpret_cast<uint32_t*>(s_b2 + C);
  for (int i = threadIdx.x; i < p.hidden / 2; i += W_THREADS) {
    uint32_t pk;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(p.b1[2 * i + 1]), "f"(p.b1[2 * i]));
    s_b1h[i] = pk;
  }
  for (int i = threadIdx.x; i < C; i += W_THREADS) s_b2[i] = p.b2[i];

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = (p.M + WPAIR_M - 1) / WPAIR_M;
  const int nch = p.hidden / WHC;
  const int my_tiles = pair_id < num_tiles ? (num_tiles - pair_id + num_pairs - 1) / num_pairs : 0;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w1);
    ptx::prefetch_tmap(&tmap_w2);
  }
  if (warp_idx == 1 && lane == 0) {
    // "leader" barriers collect one arrival per CTA (TMA) or per epilogue warp of the pair; the others are signalled
    // by the leader's multicast tcgen05.commit
    for (int s = 0; s < R1; ++s) {
      ptx::mbar_init(w1full(s), 2);
      ptx::mbar_init(w1empty(s), 1);
    }
    for (int s = 0; s < R2; ++s) {
      ptx::mbar_init(w2full(s), 2);
      ptx::mbar_init(w2empty(s), 1);
    }
    ptx::mbar_init(xfull, 2);
    ptx::mbar_init(xempty, 1);
    ptx::mbar_init(sfull, 1);
    ptx::mbar_init(sempty, 2 * W_EPI_WARPS);
    for (uint32_t b = 0; b < 2; ++b) {
      ptx::mbar_init(hfull(b), 2 * W_EPI_WARPS);
      ptx::mbar_init(hempty(b), 1);
    }
    ptx::mbar_init(ofull, 1);
    ptx::mbar_init(oempty, 2 * W_EPI_WARPS);
    ptx::fence_barrier_init();
  }
  if (warp_idx == 2) {
    ptx::tmem_alloc_pair(tmem_ptr_smem, 512);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  // under the previous kernel's tail: the pairs share the (constant) weight chunks out and pull them into L2
  if (warp_idx == 0 && lane == 0) {
    for (int j = pair_id; j < nch; j += num_pairs) {
#pragma unroll
      for (int kp = 0; kp < KP; ++kp)
        ptx::tma_prefetch_2d(&tmap_w1, kp * 64, j * WHC + static_cast<int>(cta_rank) * (WHC / 2));
#pragma unroll
      for (int nh = 0; nh < NH; ++nh) ptx::tma_prefetch_2d(&tmap_w2, j * WHC, nh * 192 + static_cast<int>(cta_rank) * 96);
    }
  }
  pdl_sync();

  if (warp_idx == 0) {
    // ===================== TMA producer: X tiles and W1 chunks (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t xfull_leader = ptx::mapa_rank(xfull, 0);
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int tile = pair_id + tl * num_pairs;
        const int m0 = tile * WPAIR_M + static_cast<int>(cta_rank) * WM;
        ptx::mbar_wait(xempty, (static_cast<uint32_t>(tl) & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx_cluster(xfull_leader, Cfg::X_BYTES);
#pragma unroll
        for (int kp = 0; kp < KP; ++kp) ptx::tma_load_2d_pair(smem_x + kp * WX_PANEL, &tmap_x, kp * 64, m0, xfull_leader);
        if (tl + 1 < my_tiles) {   // X is single-buffered: make the next tile's load an L2 hit
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) ptx::tma_prefetch_2d(&tmap_x, kp * 64, m0 + num_pairs * WPAIR_M);
        }
        for (int j = 0; j < nch; ++j) {
          ptx::mbar_wait(w1empty(stage), phase ^ 1u);
          const uint32_t full_leader = ptx::mapa_rank(w1full(stage), 0);
          ptx::mbar_arrive_expect_tx_cluster(full_leader, Cfg::W1_STAGE);
#pragma unroll
          for (int kp = 0; kp < KP; ++kp)
            ptx::tma_load_2d_pair(smem_w1 + stage * Cfg::W1_STAGE + kp * Cfg::W1_PANEL, &tmap_w1, kp * 64,
                                  j * WHC + static_cast<int>(cta_rank) * (WHC / 2), full_leader);
          if (++stage == R1) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 2) {
    // ===================== TMA producer: W2 chunks (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        for (int j = 0; j < nch; ++j) {
          ptx::mbar_wait(w2empty(stage), phase ^ 1u);
          const uint32_t full_leader = ptx::mapa_rank(w2full(stage), 0);
          ptx::mbar_arrive_expect_tx_cluster(full_leader, Cfg::W2_STAGE);
#pragma unroll
          for (int nh = 0; nh < NH; ++nh)
            ptx::tma_load_2d_pair(smem_w2 + stage * Cfg::W2_STAGE + nh * Cfg::W2_HALF, &tmap_w2, j * WHC,
                                  nh * 192 + static_cast<int>(cta_rank) * 96, full_leader);
          if (++stage == R2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== GEMM1 issuer (leader CTA only): S = X . W1[chunk]^T =====================
    // Two issuing warps (this one and warp 3): every mbarrier wait costs the waiting warp 100-150 cycles even when the
    // phase has long completed, and tcgen05.mma issue blocks while the pipe's queue is full; in one instruction stream
    // the four waits per chunk left the tensor pipe idle a third of the time (first version of this kernel: 970 TFLOP/s).
    // The whole warp runs the control flow so that descriptors stay in uniform registers; only the tcgen05
    // instructions sit under elect_one.
    if (cta_rank == 0) {
      constexpr uint32_t idesc1 = ptx::make_idesc_bf16(WPAIR_M, WHC);
      int st1 = 0;
      uint32_t ph1 = 0, gs = 0;
      const uint32_t tmem_s = tmem_base + Cfg::TMEM_S;
      for (int tl = 0; tl < my_tiles; ++tl) {
        ptx::mbar_wait(xfull, static_cast<uint32_t>(tl) & 1u);
        for (int j = 0; j < nch; ++j, ++gs) {
          ptx::mbar_wait(w1full(st1), ph1);
          ptx::mbar_wait(sempty, (gs & 1u) ^ 1u);   // the epilogue has pulled the previous S into registers
          ptx::tc_fence_after();
#pragma unroll
          for (int kp = 0; kp < KP; ++kp) {
            const uint64_t da = ptx::make_kmajor_sw128_desc(smem_x + kp * WX_PANEL);
            const uint64_t db = ptx::make_kmajor_sw128_desc(smem_w1 + st1 * Cfg::W1_STAGE + kp * Cfg::W1_PANEL);
            if (!(p.dbg & 1) && ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_bf16_pair(tmem_s, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k), idesc1,
                                    (kp | k) != 0 ? 1u : 0u);
            }
            __syncwarp();
          }
          if (ptx::elect_one()) {
            ptx::umma_commit_pair(w1empty(st1), 3);
            ptx::umma_commit_pair(sfull, 3);
            if (j == nch - 1) ptx::umma_commit_pair(xempty, 3);   // X may be refilled for the next tile
          }
          __syncwarp();
          if (++st1 == R1) { st1 = 0; ph1 ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 3) {
    // ===================== GEMM2 issuer (leader CTA only): O += H . W2[:, chunk]^T, H from tensor memory =====================
    if (cta_rank == 0) {
      constexpr uint32_t idesc2 = ptx::make_idesc_f16(WPAIR_M, 192);  // H and W2 are fp16
      int st2 = 0;
      uint32_t ph2 = 0, gh = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        for (int j = 0; j < nch; ++j, ++gh) {
          const uint32_t b = gh & 1u;
          ptx::mbar_wait(w2full(st2), ph2);
          if (j == 0 && tl > 0) ptx::mbar_wait(oempty, static_cast<uint32_t>(tl - 1) & 1u);  // O of the previous tile is in registers
          ptx::mbar_wait(hfull(b), (gh >> 1) & 1u);
          ptx::tc_fence_after();
          const uint32_t tmem_h = tmem_base + Cfg::TMEM_H + b * (WHC / 2);
#pragma unroll
          for (int nh = 0; nh < NH; ++nh) {
            const uint64_t db = ptx::make_kmajor_sw128_desc(smem_w2 + st2 * Cfg::W2_STAGE + nh * Cfg::W2_HALF);
            if (!(p.dbg & 2) && ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_f16_pair_ts(tmem_base + static_cast<uint32_t>(nh * 192), tmem_h + static_cast<uint32_t>(k * 8),
                                      db + static_cast<uint64_t>(2 * k), idesc2, (j > 0 || k != 0) ? 1u : 0u);
            }
            __syncwarp();
          }
          if (ptx::elect_one()) {
            ptx::umma_commit_pair(w2empty(st2), 3);
            ptx::umma_commit_pair(hempty(b), 3);
            if (j == nch - 1) ptx::umma_commit_pair(ofull, 3);
          }
          __syncwarp();
          if (++st2 == R2) { st2 = 0; ph2 ^= 1u; }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue (both CTAs, 16 warps each) =====================
    const int ew = warp_idx & 3;            // TMEM lane quarter == scheduler
    const int grp = (warp_idx - 4) >> 2;    // which 16 of a chunk's 64 columns / which 32-column output blocks
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const uint32_t sempty_leader = ptx::mapa_rank(sempty, 0);
    const uint32_t hfull_leader0 = ptx::mapa_rank(hfull(0), 0), hfull_leader1 = ptx::mapa_rank(hfull(1), 0);
    const uint32_t oempty_leader = ptx::mapa_rank(oempty, 0);
    uint32_t g = 0;
    for (int tl = 0; tl < my_tiles; ++tl) {
      const int tile = pair_id + tl * num_pairs;
      const int m = tile * WPAIR_M + static_cast<int>(cta_rank) * WM + ew * 32 + lane;   // this lane's row
      for (int j = 0; j < nch; ++j, ++g) {
        const uint32_t b = g & 1u, use = (g >> 1) & 1u;
        ptx::mbar_wait(sfull, g & 1u);
        ptx::tc_fence_after();
        uint32_t r[16];
        ptx::tmem_ld_32x16(lane_base + Cfg::TMEM_S + static_cast<uint32_t>(grp * 16), r);
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(sempty_leader);   // S is in registers: GEMM1 of the next chunk may run
        const uint32_t* b1h = s_b1h + (j * WHC + grp * 16) / 2;
        uint32_t hq[8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const uint4 b4 = reinterpret_cast<const uint4*>(b1h)[q];
          hq[4 * q] = gelu_half_f16x2_b(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1]), b4.x);
          hq[4 * q + 1] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3]), b4.y);
          hq[4 * q + 2] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5]), b4.z);
          hq[4 * q + 3] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7]), b4.w);
        }
        ptx::mbar_wait(hempty(b), use ^ 1u);   // GEMM2 two chunks ago has finished reading H[b]
        ptx::tc_fence_after();
        ptx::tmem_st_32x8(lane_base + Cfg::TMEM_H + b * (WHC / 2) + static_cast<uint32_t>(grp * 8), hq);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(b ? hfull_leader1 : hfull_leader0);
      }
      // ---- drain O: +b2, +resid, bf16 -> global; this warp's 32-column blocks: grp, grp + 4, grp + 8 ----
      // The residual rows are inputs: the first block's are requested before the wait for O, each next block's while
      // the current one is converted.  Each lane moves 64 contiguous bytes per block (its own row): simple, but a
      // warp-wide access touches 32 different lines, and the drain costs 78 of the launch's 313 us at M = 131072
      // (FVLA_FFN_WIDE_DEBUG=4 skips it: 235 us).  Holding the whole residual in registers and releasing O before the
      // stores measured slower (345 us): the cost is the access pattern, not its latency — the next step is to stage
      // the blocks through shared memory with TMA on both sides, as the other tcgen05 kernels do.
      const bool row_ok = m < p.M;
      const __nv_bfloat16* rrow = p.resid + static_cast<size_t>(row_ok ? m : 0) * C;
      __nv_bfloat16* orow = p.out + static_cast<size_t>(row_ok ? m : 0) * C;
      uint4 cur[4], nxt[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        cur[c] = make_uint4(0u, 0u, 0u, 0u);
        nxt[c] = make_uint4(0u, 0u, 0u, 0u);
        if (row_ok && !(p.dbg & 4)) cur[c] = *reinterpret_cast<const uint4*>(rrow + grp * 32 + c * 8);
      }
      ptx::mbar_wait(ofull, static_cast<uint32_t>(tl) & 1u);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int blk = grp; blk < ((p.dbg & 4) ? 0 : C / 32); blk += 4) {
        if (row_ok && blk + 4 < C / 32) {
#pragma unroll
          for (int c = 0; c < 4; ++c) nxt[c] = *reinterpret_cast<const uint4*>(rrow + (blk + 4) * 32 + c * 8);
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[16];
          ptx::tmem_ld_32x16(lane_base + static_cast<uint32_t>(blk * 32 + half * 16), r);
          ptx::tmem_ld_wait();
          const float* b2 = s_b2 + blk * 32 + half * 16;
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint4 rv = cur[half * 2 + c];
            const uint32_t w[4] = {rv.x, rv.y, rv.z, rv.w};
            uint32_t o[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float v0 = __uint_as_float(r[8 * c + 2 * q]) + b2[8 * c + 2 * q] + __uint_as_float(w[q] << 16);
              const float v1 = __uint_as_float(r[8 * c + 2 * q + 1]) + b2[8 * c + 2 * q + 1] +
                               __uint_as_float(w[q] & 0xffff0000u);
              o[q] = pack_bf16(v0, v1);
            }
            if (row_ok) *reinterpret_cast<uint4*>(orow + blk * 32 + (half * 2 + c) * 8) = make_uint4(o[0], o[1], o[2], o[3]);
          }
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) cur[c] = nxt[c];
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(oempty_leader);   // this warp's O columns are consumed
    }
  }

  // neither CTA may exit (or free TMEM) while the peer can still signal its barriers or the leader's MMAs read
  // its shared / tensor memory
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp_idx == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int C>
int launch_ffn_wide(const FfnFusedArgs& a, cudaStream_t stream) {
  using Cfg = WideCfg<C>;
  auto kfn = ffn_wide_kernel<C>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), Cfg::SMEM_BYTES)) return rc;
  CUtensorMap tx, tw1, tw2;
  if (int rc = make_tmap_bf16(&tx, a.x, a.M, C, C, WM)) return rc;
  if (int rc = make_tmap_bf16(&tw1, a.w1, a.hidden, C, C, WHC / 2)) return rc;
  if (int rc = make_tmap_bf16(&tw2, a.w2, C, a.hidden, a.hidden, 96)) return rc;   // fp16 bits, same geometry
  WideParams p;
  p.M = a.M; p.hidden = a.hidden; p.b1 = a.b1; p.b2 = a.b2;
  p.resid = static_cast<const __nv_bfloat16*>(a.resid);
  p.out = static_cast<__nv_bfloat16*>(a.out);
  static const int dbg = std::getenv("FVLA_FFN_WIDE_DEBUG") ? std::atoi(std::getenv("FVLA_FFN_WIDE_DEBUG")) : 0;
  p.dbg = dbg;
  const int tiles = ceil_div(a.M, WPAIR_M);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(W_THREADS), Cfg::SMEM_BYTES, stream, tx, tw1, tw2, p));
  return 0;
}

}  // namespace

bool ffn_wide_supported(int dtype, int C, int hidden) {
  static const bool on = std::getenv("FVLA_ENABLE_FFN_WIDE") != nullptr;   // opt-in until it has earned its place
  return on && dtype == DT_BF16 && C == 384 && hidden % WHC == 0 && hidden >= WHC && hidden <= 4 * C;
}

int ffn_wide(const FfnFusedArgs& a, cudaStream_t stream) {
  FVLA_REQUIRE(a.M > 0 && a.C == 384 && a.hidden % WHC == 0 && a.hidden >= WHC && a.hidden <= 4 * a.C,
               "ffn_wide: unsupported shape");
  FVLA_REQUIRE(a.b1 != nullptr && a.b2 != nullptr && a.resid != nullptr, "ffn_wide: biases and residual required");
  return launch_ffn_wide<384>(a, stream);
}

}  // namespace fvla
