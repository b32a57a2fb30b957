// Fused ConvFFN tail for the wide FastViTHD stages:   out = resid + fc2( GELU( fc1(z) ) )
// (reference: HF-hub FastViTHD `convffn.fc1 -> act -> fc2`, layer scale folded into fc2).
//
// Unfused, the 4x hidden tensor of a stage-0/1 block (M x 4C bf16: 805 MB per 32 images) is written by the fc1
// GEMM and read back by the fc2 GEMM; the fc1 launch is bound by that write (profiles/), not by its math.
// Here the hidden activations never leave the SM pair:
//
//   per 256-row tile (CTA pair, cta_group::2; each CTA owns 128 rows):
//     X tile [128 x C] resident in shared memory (one TMA load per tile)
//     for each 128-wide hidden chunk j:
//       GEMM1  S[j&1] (TMEM, 128 x 128 fp32)  = X . W1[j]^T          W1 panels streamed through a TMA ring
//       epilogue warps: S -> +b1 -> GELU (packed-half, epilogue_math.cuh) -> fp16 H[j&1] in shared memory, in the
//                       K-major SWIZZLE_128B layout the tensor core reads its A operand from
//       GEMM2  O (TMEM, 128 x C fp32)        += H[j&1] . W2[:, j]^T   fp16 x fp16; W2 (fp16) through the same ring
//     epilogue warps: O -> +b2 -> +resid -> bf16 -> staging slab -> TMA store
//
// GEMM1 runs two chunks ahead of GEMM2 (S is released once it sits in registers), so the tensor core works on
// the coming chunks while the 16 epilogue warps evaluate the GELU of this one (two S buffers, two H buffers).  TMEM: O at column 0 (C <= 192
// columns), S0 at 256, S1 at 384.  Each W panel is split across the pair (half the rows per CTA), so a CTA
// streams 2*C*4C bytes of weights per 128 rows — the same as the two unfused GEMMs — and moves NO hidden bytes
// through HBM: algorithmic traffic is read z + read resid + write out = 6*M*C bytes.
#include "common.cuh"
#include "epilogue_math.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tma_host.h"

#include <cstdlib>

namespace fvla {
namespace {
using namespace epi;

constexpr int FM = 128;            // rows per CTA
constexpr int FPAIR_M = 256;       // rows per CTA pair
constexpr int HC = 128;            // hidden columns per chunk
constexpr int FFN_THREADS = 640;   // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4..19 epilogue
constexpr int EPI_WARPS = 16;
constexpr int PANEL_BYTES = FM * 128;  // 128 rows x 64 bf16, SWIZZLE_128B
constexpr int TMEM_S0 = 256, TMEM_S1 = 384;

template <int C> struct FfnCfg {
  static constexpr int KP = (C + 63) / 64;                   // 64-wide K panels of X / W1
  static constexpr int X_BYTES = KP * PANEL_BYTES;
  static constexpr int W1_BYTES = (HC / 2) * 128;            // this CTA's half of a W1 panel: 64 rows
  static constexpr int W2_BYTES = (C / 2) * 128;             // this CTA's half of a W2 panel: C/2 rows
  static constexpr int STAGE_BYTES = ((W1_BYTES > W2_BYTES ? W1_BYTES : W2_BYTES) + 1023) / 1024 * 1024;
  static constexpr int STAGES = 8;
  static constexpr int H_BYTES = FM * HC * 2;                // one H buffer: two panels
  static constexpr int NOSUB = (C + 63) / 64;                // 64-column output sub-tiles
  static constexpr int BAR_BYTES = 512;
  static constexpr int MAX_HIDDEN = 4 * C;                   // b1 is kept in shared memory (ConvFFN ratio 4)
  static constexpr int BIAS_BYTES = (MAX_HIDDEN + C) * 4 + MAX_HIDDEN * 2;  // b1 fp32, b2 fp32, b1 as packed halves
  static constexpr int SMEM_BYTES = X_BYTES + STAGES * STAGE_BYTES + 2 * H_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
  static_assert(C % 32 == 0 && C <= 192, "O accumulators must fit TMEM columns [0, 256)");
  static_assert(2 * H_BYTES >= EPI_WARPS * 4096, "H buffers double as the output staging slabs");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

struct FfnParams {
  int M, hidden;
  const float* b1;              // [hidden] (pre-halved with W1: the GELU epilogue takes x/2)
  const float* b2;              // [C]
  const __nv_bfloat16* resid;   // [M, ldr], may alias the output
  int ldr;
};

// TEAMS: the 16 epilogue warps work as two teams of 8 (team = chunk parity = S/H buffer): each warp evaluates the GELU
// of 32 rows x 64 columns of every OTHER chunk, so the four warps of a scheduler are spread over two chunks in
// different phases (TMEM load / math / barrier hand-off) instead of marching through one chunk in lock step.
template <int C, bool TEAMS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FFN_THREADS, 1)
ffn_fused_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                 const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out,
                 const FfnParams p) {
  using Cfg = FfnCfg<C>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_ffn[];
  const uint32_t smem_base = (ptx::smem_u32(smem_ffn) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;
  const uint32_t smem_w = smem_x + Cfg::X_BYTES;
  const uint32_t smem_h = smem_w + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = smem_h + 2 * Cfg::H_BYTES;
  auto wfull = [&](int s) { return bar_base + 8u * s; };
  auto wempty = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t xfull = bar_base + 8u * (2 * STAGES), xempty = xfull + 8u;
  auto sfull = [&](int b) { return xempty + 8u + 8u * b; };
  auto sempty = [&](int b) { return xempty + 24u + 8u * b; };
  auto hfull = [&](int b) { return xempty + 40u + 8u * b; };
  auto hempty = [&](int b) { return xempty + 56u + 8u * b; };
  const uint32_t ofull = xempty + 72u, oempty = xempty + 80u;
  const uint32_t tmem_ptr_smem = xempty + 88u;
  // biases in shared memory: read by every epilogue warp for every chunk (ncu: the per-chunk __ldg of b1 was the
  // top long-scoreboard stall of the epilogue)
  float* s_b1 = reinterpret_cast<float*>(smem_ffn + (bar_base - ptx::smem_u32(smem_ffn)) + Cfg::BAR_BYTES);
  float* s_b2 = s_b1 + Cfg::MAX_HIDDEN;
  uint32_t* s_b1h = reinterpret_cast<uint32_t*>(s_b2 + C);  // b1 as f16x2 pairs (TEAMS epilogue adds the bias in half)
  for (int i = threadIdx.x; i < p.hidden; i += FFN_THREADS) s_b1[i] = p.b1[i];
  for (int i = threadIdx.x; i < p.hidden / 2; i += FFN_THREADS) {
    uint32_t pk;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(pk) : "f"(p.b1[2 * i + 1]), "f"(p.b1[2 * i]));
    s_b1h[i] = pk;
  }
  for (int i = threadIdx.x; i < C; i += FFN_THREADS) s_b2[i] = p.b2[i];

  const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_tiles = (p.M + FPAIR_M - 1) / FPAIR_M;
  const int nch = p.hidden / HC;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w1);
    ptx::prefetch_tmap(&tmap_w2);
    ptx::prefetch_tmap(&tmap_out);
  }
  if (warp_idx == 1 && lane == 0) {
    // "leader" barriers collect one arrival per CTA (TMA) or per epilogue warp of the pair; the others are
    // signalled by the leader's multicast tcgen05.commit
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(wfull(s), 2);
      ptx::mbar_init(wempty(s), 1);
    }
    ptx::mbar_init(xfull, 2);
    ptx::mbar_init(xempty, 1);
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(sfull(b), 1);
      ptx::mbar_init(sempty(b), TEAMS ? EPI_WARPS : 2 * EPI_WARPS);
      ptx::mbar_init(hfull(b), TEAMS ? EPI_WARPS : 2 * EPI_WARPS);
      ptx::mbar_init(hempty(b), 1);
    }
    ptx::mbar_init(ofull, 1);
    ptx::mbar_init(oempty, 2 * EPI_WARPS);
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

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0, t = 0;
      const uint32_t xfull_leader = ptx::mapa_rank(xfull, 0);
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++t) {
        const int m0 = tile * FPAIR_M + static_cast<int>(cta_rank) * FM;
        ptx::mbar_wait(xempty, (t & 1u) ^ 1u);
        ptx::mbar_arrive_expect_tx_cluster(xfull_leader, Cfg::X_BYTES);
#pragma unroll
        for (int kp = 0; kp < Cfg::KP; ++kp)
          ptx::tma_load_2d_pair(smem_x + kp * PANEL_BYTES, &tmap_x, kp * 64, m0, xfull_leader);
        // weight panels in exactly the order the MMA warp consumes them:
        //   W1(0), W1(1), then for every chunk j: W1(j+2) (if any), W2(j)
        auto load_w1 = [&](int j) {
          for (int kp = 0; kp < Cfg::KP; ++kp) {
            ptx::mbar_wait(wempty(stage), phase ^ 1u);
            const uint32_t full_leader = ptx::mapa_rank(wfull(stage), 0);
            ptx::mbar_arrive_expect_tx_cluster(full_leader, Cfg::W1_BYTES);
            ptx::tma_load_2d_pair(smem_w + stage * Cfg::STAGE_BYTES, &tmap_w1, kp * 64,
                                  j * HC + static_cast<int>(cta_rank) * (HC / 2), full_leader);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        };
        auto load_w2 = [&](int j) {
          for (int kb = 0; kb < HC / 64; ++kb) {
            ptx::mbar_wait(wempty(stage), phase ^ 1u);
            const uint32_t full_leader = ptx::mapa_rank(wfull(stage), 0);
            ptx::mbar_arrive_expect_tx_cluster(full_leader, Cfg::W2_BYTES);
            ptx::tma_load_2d_pair(smem_w + stage * Cfg::STAGE_BYTES, &tmap_w2, j * HC + kb * 64,
                                  static_cast<int>(cta_rank) * (C / 2), full_leader);
            if (++stage == STAGES) { stage = 0; phase ^= 1u; }
          }
        };
        load_w1(0);
        if (nch > 1) load_w1(1);
        for (int j = 0; j < nch; ++j) {
          if (j + 2 < nch) load_w1(j + 2);
          load_w2(j);
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // The whole warp runs the control flow (barrier waits, stage arithmetic) so that descriptors and addresses
    // stay in uniform registers; only the tcgen05 instructions themselves sit under elect_one.  With a single
    // divergent thread every UTCHMMA cost ~15 dependent instructions of operand marshalling, comparable to
    // the 64 cycles an M=256, N=128 instruction executes for.
    if (cta_rank == 0) {
      constexpr uint32_t idesc1 = ptx::make_idesc_bf16(FPAIR_M, HC);
      constexpr uint32_t idesc2 = ptx::make_idesc_f16(FPAIR_M, C);  // H and W2 are fp16
      int stage = 0;
      uint32_t phase = 0, t = 0, g1 = 0, g2 = 0;
      // GEMM1 runs two chunks ahead of GEMM2: S[b] is released as soon as the epilogue has pulled it into
      // registers, so GEMM1(j+2) executes while GELU(j) is being evaluated and the epilogue never waits for S;
      // GEMM2(j) follows whenever H[j] is ready.  (With GEMM1 only one chunk ahead and S released at the end of
      // the GELU, every chunk paid two barrier round trips plus both GEMMs in series: tensor pipe 37 % busy.)
      auto gemm1 = [&](int j) {
        const uint32_t b = g1 & 1u;
        ptx::mbar_wait(sempty(b), ((g1 >> 1) & 1u) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_s = tmem_base + (b ? TMEM_S1 : TMEM_S0);
#pragma unroll
        for (int kp = 0; kp < Cfg::KP; ++kp) {
          ptx::mbar_wait(wfull(stage), phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_kmajor_sw128_desc(smem_x + kp * PANEL_BYTES);
          const uint64_t db = ptx::make_kmajor_sw128_desc(smem_w + stage * Cfg::STAGE_BYTES);
          constexpr int NK_FULL = 4;
          const int nk = (C - kp * 64) / 16 < NK_FULL ? (C - kp * 64) / 16 : NK_FULL;
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < NK_FULL; ++k)
              if (k < nk)
                ptx::umma_bf16_pair(tmem_s, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                                    idesc1, (kp | k) != 0 ? 1u : 0u);
            ptx::umma_commit_pair(wempty(stage), 3);
            if (kp == Cfg::KP - 1) {
              ptx::umma_commit_pair(sfull(b), 3);
              if (j == nch - 1) ptx::umma_commit_pair(xempty, 3);  // X may be refilled for the next tile
            }
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ++g1;
      };
      auto gemm2 = [&](int j, uint32_t t) {
        const uint32_t b = g2 & 1u;
        if (j == 0) {  // O of the previous tile must have been drained
          ptx::mbar_wait(oempty, (t & 1u) ^ 1u);
          ptx::tc_fence_after();
        }
        ptx::mbar_wait(hfull(b), (g2 >> 1) & 1u);
        ptx::tc_fence_after();
#pragma unroll
        for (int kb = 0; kb < HC / 64; ++kb) {
          ptx::mbar_wait(wfull(stage), phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_kmajor_sw128_desc(smem_h + b * Cfg::H_BYTES + kb * PANEL_BYTES);
          const uint64_t db = ptx::make_kmajor_sw128_desc(smem_w + stage * Cfg::STAGE_BYTES);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              ptx::umma_bf16_pair(tmem_base, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                                  idesc2, (j > 0 || kb != 0 || k != 0) ? 1u : 0u);
            ptx::umma_commit_pair(wempty(stage), 3);
            if (kb == HC / 64 - 1) ptx::umma_commit_pair(hempty(b), 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        ++g2;
      };
      for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++t) {
        ptx::mbar_wait(xfull, t & 1u);
        ptx::tc_fence_after();
        gemm1(0);
        if (nch > 1) gemm1(1);
        for (int j = 0; j < nch; ++j) {
          if (j + 2 < nch) gemm1(j + 2);
          gemm2(j, t);
        }
        if (ptx::elect_one()) ptx::umma_commit_pair(ofull, 3);
        __syncwarp();
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue (both CTAs, 16 warps each) =====================
    const int ew = warp_idx & 3;            // TMEM lane quarter == scheduler
    const int grp = (warp_idx - 4) >> 2;    // which 32 of a chunk's 128 columns / which output sub-tile
    const int row = ew * 32 + lane;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const uint32_t sempty_leader0 = ptx::mapa_rank(sempty(0), 0), sempty_leader1 = ptx::mapa_rank(sempty(1), 0);
    const uint32_t hfull_leader0 = ptx::mapa_rank(hfull(0), 0), hfull_leader1 = ptx::mapa_rank(hfull(1), 0);
    const uint32_t oempty_leader = ptx::mapa_rank(oempty, 0);
    const uint32_t slab = smem_h + static_cast<uint32_t>(warp_idx - 4) * 4096u;  // output staging (H is idle then)
    uint32_t g = 0, t = 0;
    for (int tile = pair_id; tile < num_tiles; tile += num_pairs, ++t) {
      const int m0 = tile * FPAIR_M + static_cast<int>(cta_rank) * FM;
      if constexpr (TEAMS) {
        const uint32_t team = static_cast<uint32_t>(grp >> 1), half = static_cast<uint32_t>(grp & 1);
        for (int j = 0; j < nch; ++j, ++g) {
          const uint32_t b = g & 1u, use = (g >> 1) & 1u;
          if (b != team) continue;
          ptx::mbar_wait(sfull(b), use);
          ptx::tc_fence_after();
          const uint32_t hrow = smem_h + b * Cfg::H_BYTES + half * PANEL_BYTES + row * 128;
          const uint32_t* b1h = s_b1h + (j * HC + half * 64) / 2;
#pragma unroll
          for (int hp = 0; hp < 2; ++hp) {
            uint32_t r[32];
            ptx::tmem_ld_32x32(lane_base + (b ? TMEM_S1 : TMEM_S0) + half * 64u + static_cast<uint32_t>(hp * 32), r);
            ptx::tmem_ld_wait();
            if (hp == 1) {
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster(b ? sempty_leader1 : sempty_leader0);  // S[b] is in registers
            } else {
              ptx::mbar_wait(hempty(b), use ^ 1u);  // GEMM2 two chunks ago has finished reading H[b]
            }
            uint32_t hq[16];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 b4 = reinterpret_cast<const uint4*>(b1h + hp * 16)[q];
              hq[4 * q] = gelu_half_f16x2_b(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1]), b4.x);
              hq[4 * q + 1] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3]), b4.y);
              hq[4 * q + 2] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5]), b4.z);
              hq[4 * q + 3] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7]), b4.w);
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int chunk = hp * 4 + c;
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hrow + ((chunk ^ (row & 7)) << 4)),
                           "r"(hq[4 * c]), "r"(hq[4 * c + 1]), "r"(hq[4 * c + 2]), "r"(hq[4 * c + 3])
                           : "memory");
            }
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();  // generic-proxy writes of H -> visible to the tensor core's async reads
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(b ? hfull_leader1 : hfull_leader0);
        }
      } else {
      for (int j = 0; j < nch; ++j, ++g) {
          const uint32_t b = g & 1u, use = (g >> 1) & 1u;
          ptx::mbar_wait(sfull(b), use);
          ptx::tc_fence_after();
          uint32_t r[32];
          ptx::tmem_ld_32x32(lane_base + (b ? TMEM_S1 : TMEM_S0) + static_cast<uint32_t>(grp * 32), r);
          ptx::tmem_ld_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(b ? sempty_leader1 : sempty_leader0);  // S[b] is in registers
          ptx::mbar_wait(hempty(b), use ^ 1u);  // GEMM2 two chunks ago has finished reading H[b]
          const float* b1 = s_b1 + j * HC + grp * 32;
          uint32_t hq[16];  // 32 GELU outputs as fp16 pairs
  #pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = reinterpret_cast<const float4*>(b1)[q];
            hq[2 * q] = gelu_half_f16x2(__uint_as_float(r[4 * q]) + b4.x, __uint_as_float(r[4 * q + 1]) + b4.y);
            hq[2 * q + 1] = gelu_half_f16x2(__uint_as_float(r[4 * q + 2]) + b4.z, __uint_as_float(r[4 * q + 3]) + b4.w);
          }
          // H[b]: two K-major panels of 64 columns; this warp's 32 columns are chunks 4*(grp&1) .. +3 of panel grp>>1
          const uint32_t hrow = smem_h + b * Cfg::H_BYTES + static_cast<uint32_t>(grp >> 1) * PANEL_BYTES + row * 128;
  #pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int chunk = (grp & 1) * 4 + c;
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(hrow + ((chunk ^ (row & 7)) << 4)),
                         "r"(hq[4 * c]), "r"(hq[4 * c + 1]), "r"(hq[4 * c + 2]), "r"(hq[4 * c + 3])
                         : "memory");
          }
          ptx::tc_fence_before();
          ptx::fence_proxy_async_smem();  // generic-proxy writes of H -> visible to the tensor core's async reads
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(b ? hfull_leader1 : hfull_leader0);
        }
      }
      // ---- drain O: +b2, +resid, bf16, TMA store (one 64-column sub-tile per warp group; NOSUB <= 3) ----
      const int sub = grp;
      const bool has_sub = sub < Cfg::NOSUB;
      const int col0 = sub * 64;
      uint4 rr[8];
      if (has_sub) {  // residual sub-tile (32 rows x 128 B): issued before the wait for O so its latency overlaps
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = i * 32 + lane;
          const int rrow = idx >> 3, rchunk = idx & 7;
          const int gm = m0 + ew * 32 + rrow, gc = col0 + rchunk * 8;
          rr[i] = make_uint4(0u, 0u, 0u, 0u);
          if (gm < p.M && gc + 8 <= C)
            rr[i] = __ldg(reinterpret_cast<const uint4*>(p.resid + static_cast<size_t>(gm) * p.ldr + gc));
        }
      }
      ptx::mbar_wait(ofull, t & 1u);
      ptx::tc_fence_after();
      if (has_sub) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int idx = i * 32 + lane;
          const int rrow = idx >> 3, rchunk = idx & 7;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(slab + rrow * 128 + ((rchunk ^ (rrow & 7)) << 4)),
                       "r"(rr[i].x), "r"(rr[i].y), "r"(rr[i].z), "r"(rr[i].w)
                       : "memory");
        }
        __syncwarp();
        const uint32_t sbase = slab + lane * 128;
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          const int nb = col0 + hp * 32;
          if (nb >= C) break;
          uint32_t r[32];
          ptx::tmem_ld_32x32(lane_base + static_cast<uint32_t>(nb), r);
          ptx::tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 b4 = reinterpret_cast<const float4*>(s_b2 + nb)[q];
            v[4 * q] = __uint_as_float(r[4 * q]) + b4.x;
            v[4 * q + 1] = __uint_as_float(r[4 * q + 1]) + b4.y;
            v[4 * q + 2] = __uint_as_float(r[4 * q + 2]) + b4.z;
            v[4 * q + 3] = __uint_as_float(r[4 * q + 3]) + b4.w;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int chunk = hp * 4 + c;
            const uint32_t dst = sbase + ((chunk ^ (lane & 7)) << 4);
            uint32_t w0, w1, w2, w3;
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(dst));
            const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              v[8 * c + 2 * q] += __uint_as_float(w[q] << 16);
              v[8 * c + 2 * q + 1] += __uint_as_float(w[q] & 0xffff0000u);
            }
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(v[8 * c], v[8 * c + 1])),
                         "r"(pack_bf16(v[8 * c + 2], v[8 * c + 3])), "r"(pack_bf16(v[8 * c + 4], v[8 * c + 5])),
                         "r"(pack_bf16(v[8 * c + 6], v[8 * c + 7]))
                         : "memory");
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          ptx::tma_store_2d(&tmap_out, col0, m0 + ew * 32, slab);
          ptx::tma_store_commit();
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        ptx::mbar_arrive_cluster(oempty_leader);
        ptx::tma_store_wait_read<0>();  // the slabs live in H: drained before the next tile's GELU writes H
      }
      asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
    }
    if (lane == 0) ptx::tma_store_wait<0>();
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp_idx == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int C, bool TEAMS>
int launch_ffn(const FfnFusedArgs& a, cudaStream_t stream) {
  using Cfg = FfnCfg<C>;
  auto kfn = ffn_fused_kernel<C, TEAMS>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), Cfg::SMEM_BYTES)) return rc;
  CUtensorMap tx, tw1, tw2, to;
  if (int rc = make_tmap_bf16(&tx, a.x, a.M, C, C, FM)) return rc;
  if (int rc = make_tmap_bf16(&tw1, a.w1, a.hidden, C, C, HC / 2)) return rc;
  if (int rc = make_tmap_bf16(&tw2, a.w2, C, a.hidden, a.hidden, C / 2)) return rc;  // fp16 bits, same geometry
  if (int rc = make_tmap_bf16(&to, a.out, a.M, C, C, 32)) return rc;
  FfnParams p;
  p.M = a.M; p.hidden = a.hidden; p.b1 = a.b1; p.b2 = a.b2;
  p.resid = static_cast<const __nv_bfloat16*>(a.resid); p.ldr = C;
  const int tiles = ceil_div(a.M, FPAIR_M);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  kfn<<<grid, FFN_THREADS, Cfg::SMEM_BYTES, stream>>>(tx, tw1, tw2, to, p);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

bool ffn_fused_supported(int dtype, int C, int hidden) {
  return dtype == DT_BF16 && (C == 96 || C == 192) && hidden % HC == 0 && hidden >= HC && hidden <= 4 * C;
}

int ffn_fused(const FfnFusedArgs& a, cudaStream_t stream) {
  FVLA_REQUIRE(a.M > 0 && ffn_fused_supported(DT_BF16, a.C, a.hidden), "ffn_fused: unsupported shape");
  FVLA_REQUIRE(a.b1 != nullptr && a.b2 != nullptr && a.resid != nullptr, "ffn_fused: biases and residual required");
  static const bool teams = std::getenv("FVLA_FFN_NO_TEAMS") == nullptr;  // A/B switch for profiling
  if (a.C == 96) return teams ? launch_ffn<96, true>(a, stream) : launch_ffn<96, false>(a, stream);
  return teams ? launch_ffn<192, true>(a, stream) : launch_ffn<192, false>(a, stream);
}

}  // namespace fvla
