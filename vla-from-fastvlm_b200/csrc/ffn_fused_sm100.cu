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

#include <cstdio>
#include <cstdlib>

namespace fvla {
namespace {
using namespace epi;
#ifdef FVLA_FFN_TRACE_BUILD
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
#endif

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
  unsigned long long* trace;    // FVLA_FFN_TRACE: cycles the MMA warp of pair 0 spent in each wait (debug only)
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
  pdl_sync();  // prologue above (biases, barriers, TMEM) is input-independent

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


// ------------------------------------------------------------------------------------------------------------------
// H in TENSOR MEMORY.  ncu of the kernel above: the tensor pipe is busy 53-60 % of the time with no dominant stall —
// the limiter is shared-memory bandwidth.  Per 128-wide hidden chunk a CTA moves 72 KB of GEMM1 operands (the X tile
// is re-read for every chunk), 56 KB of GEMM2 operands, 32 KB of H written by the epilogue and 48 KB of weight panels
// landed by TMA: 208 KB per 1536 tensor-pipe cycles = 135 B/clk against the 128 B/clk a SM's shared memory delivers.
// Here the GELU output goes registers -> tensor memory (tcgen05.st, fp16 pairs = one 32-bit column per K pair) and
// GEMM2 takes it as its A operand from there (tcgen05.mma ..., [tmem_a], b_desc): GEMM2 reads only its W2 panels from
// shared memory and the H stores leave it entirely — 144 KB per chunk, 94 B/clk.
//   TMEM columns: O [0, C)   H [192, 256) (single buffer: 128 fp16 columns)   S0 [256, 384)   S1 [384, 512)
// H is handed over per 64-column half (GEMM2's first four MMAs read half 0): the epilogue of chunk j+1 has its values
// packed in registers long before GEMM2(j) retires and stores them the moment its half is free.  The residual tile is
// landed in the output staging slab by TMA (no registers held across the wait for O), and O is released as soon as a
// warp's accumulator columns are in registers, before the bias/residual/pack/store work.
template <int C> struct FfnTCfg {
  static constexpr int KP = (C + 63) / 64;
  static constexpr int X_BYTES = KP * PANEL_BYTES;
  static constexpr int W1_PANEL = (HC / 2) * 128;            // this CTA's half of a W1 panel: 64 rows x 128 B
  static constexpr int W2_PANEL = ((C / 2) * 128 + 1023) / 1024 * 1024;   // this CTA's half of a W2 panel: C/2 rows x 128 B
  static constexpr int W1_STAGE = KP * W1_PANEL;             // one stage = the W1 rows of a whole 128-wide hidden chunk
  static constexpr int W2_STAGE = (HC / 64) * W2_PANEL;      // one stage = the W2 columns of a whole chunk
  static constexpr int W1_TX = KP * W1_PANEL, W2_TX = (HC / 64) * (C / 2) * 128;
  static constexpr int R1 = C > 96 ? 3 : 4;                  // ring depths, in chunks
  static constexpr int R2 = C > 96 ? 2 : 4;
  static constexpr int NOSUB = (C + 63) / 64;
  static constexpr int SLAB_BYTES = 4 * NOSUB * 4096;        // one 32-row x 64-column staging slab per draining warp
  static constexpr int BAR_BYTES = 512;
  static constexpr int MAX_HIDDEN = 4 * C;
  static constexpr int BIAS_BYTES = C * 4 + MAX_HIDDEN * 2;  // b2 fp32, b1 as packed halves
  static constexpr int SMEM_BYTES = X_BYTES + R1 * W1_STAGE + R2 * W2_STAGE + SLAB_BYTES + BAR_BYTES + BIAS_BYTES + 1024;
  static constexpr int TMEM_H = 192;
  static_assert(C % 32 == 0 && C <= 192, "O accumulators must fit TMEM columns [0, 192)");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

// Scheduling.  An event trace of the single-issuer version (clock64 stamps from the issuing warp and two epilogue
// warps) showed the ISSUING WARP as the bottleneck, not the tensor pipe or the epilogue: every mbarrier wait costs it
// 100-150 cycles even when the phase has long completed, tcgen05.mma issue blocks while the pipe's short queue is full,
// and with seven waits per chunk (S buffer, three W1 panels, H, two W2 panels) in one instruction stream the pipe sat
// idle ~1000 of every 2500 cycles.  Hence:
//   * TWO issuing warps in the leader CTA — warp 1 issues GEMM1 (X . W1^T -> S), warp 3 issues GEMM2 (H . W2^T -> O).
//     The two streams only meet through the epilogue (S -> GELU -> H), so each warp's barrier latency is hidden behind
//     the other warp's MMAs;
//   * W1 and W2 stream through separate TMA rings whose stages hold a whole chunk (one wait per chunk and GEMM);
//   * the chunk stream is flat across tiles: GEMM1 of the next tile's first chunks runs under the tail of this tile.
template <int C>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FFN_THREADS, 1)
ffn_fused_tmemh_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                       const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out,
                       const __grid_constant__ CUtensorMap tmap_res, const FfnParams p) {
  using Cfg = FfnTCfg<C>;
  constexpr int R1 = Cfg::R1, R2 = Cfg::R2;
  constexpr uint32_t TMEM_H = Cfg::TMEM_H;
  extern __shared__ uint8_t smem_ffn[];
  const uint32_t smem_base = (ptx::smem_u32(smem_ffn) + 1023u) & ~1023u;
  const uint32_t smem_x = smem_base;
  const uint32_t smem_w1 = smem_x + Cfg::X_BYTES;
  const uint32_t smem_w2 = smem_w1 + R1 * Cfg::W1_STAGE;
  const uint32_t smem_slab = smem_w2 + R2 * Cfg::W2_STAGE;
  const uint32_t bar_base = smem_slab + Cfg::SLAB_BYTES;
  auto w1full = [&](int s) { return bar_base + 8u * s; };
  auto w1empty = [&](int s) { return bar_base + 8u * (R1 + s); };
  auto w2full = [&](int s) { return bar_base + 8u * (2 * R1 + s); };
  auto w2empty = [&](int s) { return bar_base + 8u * (2 * R1 + R2 + s); };
  const uint32_t xfull = bar_base + 8u * (2 * R1 + 2 * R2), xempty = xfull + 8u;
  auto sfull = [&](uint32_t b) { return xempty + 8u + 8u * b; };
  auto sempty = [&](uint32_t b) { return xempty + 24u + 8u * b; };
  auto hfull = [&](uint32_t h) { return xempty + 40u + 8u * h; };
  // hempty(chunk parity, half): GEMM2 of a chunk of that parity has finished reading that 64-column half of H
  auto hempty = [&](uint32_t par, uint32_t h) { return xempty + 56u + 8u * (2u * par + h); };
  const uint32_t ofull = xempty + 88u, oempty = xempty + 96u;
  auto rbar = [&](int w) { return xempty + 104u + 8u * w; };   // per draining warp: its residual tile has landed
  const uint32_t tmem_ptr_smem = xempty + 104u + 8u * EPI_WARPS;
  static_assert(8 * (2 * R1 + 2 * R2 + 1) + 104 + 8 * EPI_WARPS + 4 <= Cfg::BAR_BYTES, "barrier block too small");
  float* s_b2 = reinterpret_cast<float*>(smem_ffn + (bar_base - ptx::smem_u32(smem_ffn)) + Cfg::BAR_BYTES);
  uint32_t* s_b1h = reinterpret_cast<uint32_t*>(s_b2 + C);
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
  const int my_tiles = pair_id < num_tiles ? (num_tiles - pair_id + num_pairs - 1) / num_pairs : 0;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_w1);
    ptx::prefetch_tmap(&tmap_w2);
    ptx::prefetch_tmap(&tmap_out);
    ptx::prefetch_tmap(&tmap_res);
  }
  if (warp_idx == 1 && lane == 0) {
    // "full" barriers of the leader collect one arrival per CTA (TMA) or per epilogue warp of the pair; the "empty"
    // ones are signalled in both CTAs by the leader's multicast tcgen05.commit
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
    for (uint32_t b = 0; b < 2; ++b) {
      ptx::mbar_init(sfull(b), 1);
      ptx::mbar_init(sempty(b), EPI_WARPS);     // the 8 warps of a team, in both CTAs
      ptx::mbar_init(hfull(b), EPI_WARPS / 2);  // the 4 warps of a team that own this half, in both CTAs
      ptx::mbar_init(hempty(b, 0), 1);
      ptx::mbar_init(hempty(b, 1), 1);
    }
    ptx::mbar_init(ofull, 1);
    ptx::mbar_init(oempty, 2 * 4 * Cfg::NOSUB);
    for (int w = 0; w < EPI_WARPS; ++w) ptx::mbar_init(rbar(w), 1);
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
  // under the previous kernel's tail: the pairs share the (constant) weight panels out, one hidden chunk each, and
  // pull them into L2 — at small M the weights are most of this launch's HBM traffic
  if (warp_idx == 0 && lane == 0) {
    for (int j = pair_id; j < nch; j += num_pairs) {
#pragma unroll
      for (int kp = 0; kp < Cfg::KP; ++kp)
        ptx::tma_prefetch_2d(&tmap_w1, kp * 64, j * HC + static_cast<int>(cta_rank) * (HC / 2));
#pragma unroll
      for (int kb = 0; kb < HC / 64; ++kb)
        ptx::tma_prefetch_2d(&tmap_w2, j * HC + kb * 64, static_cast<int>(cta_rank) * (C / 2));
    }
  }
  pdl_sync();  // prologue above (biases, barriers, TMEM) is input-independent
#ifdef FVLA_FFN_TRACE_BUILD
  // event log of the leader CTA of pair 0 (region 0: GEMM1 warp, 1: GEMM2 warp, 2: epilogue warp 4, 3: epilogue warp 12)
  const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && lane_id() == 0;
  int tr_n = 0;
  auto ev = [&](int region, int id) {
    if (tr_on && tr_n < 254) {
      p.trace[region * 512 + 2 * tr_n] = static_cast<unsigned long long>(id);
      p.trace[region * 512 + 2 * tr_n + 1] = static_cast<unsigned long long>(clock64());
      ++tr_n;
    }
  };
#else
  auto ev = [&](int, int) {};
#endif

  if (warp_idx == 0) {
    // ===================== TMA producer: X tiles and W1 chunks (both CTAs) =====================
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
        // X is single-buffered (shared memory is full): its load can only start once the last GEMM1 of this tile has
        // retired and then sits on the next tile's critical path (event trace: 2900 cycles from HBM).  Pull the next
        // tile's rows into L2 now, a whole tile ahead, so that load becomes an L2 hit.
        if (tile + num_pairs < num_tiles) {
#pragma unroll
          for (int kp = 0; kp < Cfg::KP; ++kp)
            ptx::tma_prefetch_2d(&tmap_x, kp * 64, m0 + num_pairs * FPAIR_M);
        }
        for (int j = 0; j < nch; ++j) {
          ptx::mbar_wait(w1empty(stage), phase ^ 1u);
          const uint32_t full_leader = ptx::mapa_rank(w1full(stage), 0);
          ptx::mbar_arrive_expect_tx_cluster(full_leader, Cfg::W1_TX);
#pragma unroll
          for (int kp = 0; kp < Cfg::KP; ++kp)
            ptx::tma_load_2d_pair(smem_w1 + stage * Cfg::W1_STAGE + kp * Cfg::W1_PANEL, &tmap_w1, kp * 64,
                                  j * HC + static_cast<int>(cta_rank) * (HC / 2), full_leader);
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
          ptx::mbar_arrive_expect_tx_cluster(full_leader, Cfg::W2_TX);
#pragma unroll
          for (int kb = 0; kb < HC / 64; ++kb)
            ptx::tma_load_2d_pair(smem_w2 + stage * Cfg::W2_STAGE + kb * Cfg::W2_PANEL, &tmap_w2, j * HC + kb * 64,
                                  static_cast<int>(cta_rank) * (C / 2), full_leader);
          if (++stage == R2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== GEMM1 issuer (leader CTA only): S[b] = X . W1[chunk]^T =====================
    // The whole warp runs the control flow so that descriptors and addresses stay in uniform registers; only the
    // tcgen05 instructions themselves sit under elect_one.
    if (cta_rank == 0) {
      constexpr uint32_t idesc1 = ptx::make_idesc_bf16(FPAIR_M, HC);
      int stage = 0;
      uint32_t phase = 0, g = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        for (int j = 0; j < nch; ++j, ++g) {
          const uint32_t b = g & 1u;
          ev(0, 1000 + static_cast<int>(g));
          ptx::mbar_wait(sempty(b), ((g >> 1) & 1u) ^ 1u);   // S[b] of two chunks ago is in the epilogue's registers
          ev(0, 1100 + static_cast<int>(g));
          if (j == 0) ptx::mbar_wait(xfull, static_cast<uint32_t>(tl) & 1u);
          ptx::mbar_wait(w1full(stage), phase);
          ptx::tc_fence_after();
          ev(0, 1200 + static_cast<int>(g));
          const uint32_t tmem_s = tmem_base + (b ? TMEM_S1 : TMEM_S0);
          if (ptx::elect_one()) {
#pragma unroll
            for (int kp = 0; kp < Cfg::KP; ++kp) {
              const uint64_t da = ptx::make_kmajor_sw128_desc(smem_x + kp * PANEL_BYTES);
              const uint64_t db = ptx::make_kmajor_sw128_desc(smem_w1 + stage * Cfg::W1_STAGE + kp * Cfg::W1_PANEL);
              constexpr int NK_FULL = 4;
              const int nk = (C - kp * 64) / 16 < NK_FULL ? (C - kp * 64) / 16 : NK_FULL;
#pragma unroll
              for (int k = 0; k < NK_FULL; ++k)
                if (k < nk)
                  ptx::umma_bf16_pair(tmem_s, da + static_cast<uint64_t>(2 * k), db + static_cast<uint64_t>(2 * k),
                                      idesc1, (kp | k) != 0 ? 1u : 0u);
            }
            ptx::umma_commit_pair(w1empty(stage), 3);
            ptx::umma_commit_pair(sfull(b), 3);
            if (j == nch - 1) ptx::umma_commit_pair(xempty, 3);  // X may be refilled for the next tile
          }
          __syncwarp();
          ev(0, 1300 + static_cast<int>(g));
          if (++stage == R1) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 3) {
    // ===================== GEMM2 issuer (leader CTA only): O += H . W2[:, chunk]^T, H from tensor memory =====================
    if (cta_rank == 0) {
      constexpr uint32_t idesc2 = ptx::make_idesc_f16(FPAIR_M, C);  // H and W2 are fp16
      int stage = 0;
      uint32_t phase = 0, g = 0;
      for (int tl = 0; tl < my_tiles; ++tl) {
        for (int j = 0; j < nch; ++j, ++g) {
          ptx::mbar_wait(w2full(stage), phase);
          if (j == 0 && tl > 0) ptx::mbar_wait(oempty, static_cast<uint32_t>(tl - 1) & 1u);  // O is in registers
#pragma unroll
          for (int kb = 0; kb < HC / 64; ++kb) {
            ev(1, 2000 + static_cast<int>(g) * 2 + kb);
            ptx::mbar_wait(hfull(static_cast<uint32_t>(kb)), g & 1u);
            ptx::tc_fence_after();
            ev(1, 2200 + static_cast<int>(g) * 2 + kb);
            const uint64_t db = ptx::make_kmajor_sw128_desc(smem_w2 + stage * Cfg::W2_STAGE + kb * Cfg::W2_PANEL);
            if (ptx::elect_one()) {
#pragma unroll
              for (int k = 0; k < 4; ++k)
                ptx::umma_f16_pair_ts(tmem_base, tmem_base + TMEM_H + static_cast<uint32_t>(kb * 32 + k * 8),
                                      db + static_cast<uint64_t>(2 * k), idesc2, (j > 0 || kb != 0 || k != 0) ? 1u : 0u);
              ptx::umma_commit_pair(hempty(g & 1u, static_cast<uint32_t>(kb)), 3);  // this half of H may be overwritten
              if (kb == HC / 64 - 1) {
                ptx::umma_commit_pair(w2empty(stage), 3);
                if (j == nch - 1) ptx::umma_commit_pair(ofull, 3);
              }
            }
            __syncwarp();
            ev(1, 2400 + static_cast<int>(g) * 2 + kb);
          }
          if (++stage == R2) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue (both CTAs, 16 warps each: two teams of 8, team = chunk parity) =====================
    const int ew = warp_idx & 3;            // TMEM lane quarter == scheduler
    const int grp = (warp_idx - 4) >> 2;
    const uint32_t team = static_cast<uint32_t>(grp >> 1), half = static_cast<uint32_t>(grp & 1);
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
    const uint32_t sempty_leader0 = ptx::mapa_rank(sempty(0), 0), sempty_leader1 = ptx::mapa_rank(sempty(1), 0);
    const uint32_t hfull_leader = ptx::mapa_rank(hfull(half), 0);
    const uint32_t oempty_leader = ptx::mapa_rank(oempty, 0);
    const int sub = grp;                    // 64-column output sub-tile drained by this warp
    const bool has_sub = sub < Cfg::NOSUB;
    const uint32_t slab = smem_slab + static_cast<uint32_t>((has_sub ? sub : 0) * 4 + ew) * 4096u;
    const uint32_t my_rbar = rbar(warp_idx - 4);
    const int col0 = sub * 64;
    auto request_resid = [&](int tile) {    // lane 0: residual rows of `tile` -> slab (once the last store has read it)
      ptx::tma_store_wait_read<0>();
      ptx::mbar_arrive_expect_tx(my_rbar, 4096);
      ptx::tma_load_2d(slab, &tmap_res, col0, tile * FPAIR_M + static_cast<int>(cta_rank) * FM + ew * 32, my_rbar);
    };
    if (has_sub && lane == 0 && pair_id < num_tiles) request_resid(pair_id);
    // ---- drain O of local tile `dl`: +b2, +resid, bf16, TMA store (one 64-column sub-tile per warp; NOSUB <= 3) ----
    int resid_tile = -1;                    // tile whose residual rows are still to be requested into the slab
    auto drain = [&](int dl) {
      const int tile = pair_id + dl * num_pairs;
      const int m0 = tile * FPAIR_M + static_cast<int>(cta_rank) * FM;
      if (resid_tile >= 0) {   // no GELU of this warp ran since the last drain (e.g. two chunks per tile): request now
        if (lane == 0) request_resid(resid_tile);
        resid_tile = -1;
        __syncwarp();
      }
      ptx::mbar_wait(my_rbar, static_cast<uint32_t>(dl) & 1u);    // residual tile in the slab (32 rows x 128 B, SWIZZLE_128B)
      ptx::mbar_wait(ofull, static_cast<uint32_t>(dl) & 1u);
      ptx::tc_fence_after();
      const uint32_t sbase = slab + lane * 128;
#pragma unroll
      for (int hp = 0; hp < 2; ++hp) {
        const int nb = col0 + hp * 32;
        if (nb >= C) break;
        uint32_t r[32];
        ptx::tmem_ld_32x32(lane_base + static_cast<uint32_t>(nb), r);
        ptx::tmem_ld_wait();
        if (hp == 1 || nb + 32 >= C) {    // the accumulator is in registers: GEMM2 of the next tile may overwrite it
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(oempty_leader);
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int chunk = hp * 4 + c;
          const uint32_t dst = sbase + ((chunk ^ (lane & 7)) << 4);
          uint32_t w0, w1, w2, w3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3) : "r"(dst));
          const uint32_t w[4] = {w0, w1, w2, w3};
          const float4 ba = reinterpret_cast<const float4*>(s_b2 + nb)[2 * c];
          const float4 bb = reinterpret_cast<const float4*>(s_b2 + nb)[2 * c + 1];
          const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          float v[8];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            v[2 * q] = __uint_as_float(r[8 * c + 2 * q]) + bias[2 * q] + __uint_as_float(w[q] << 16);
            v[2 * q + 1] = __uint_as_float(r[8 * c + 2 * q + 1]) + bias[2 * q + 1] + __uint_as_float(w[q] & 0xffff0000u);
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16(v[0], v[1])),
                       "r"(pack_bf16(v[2], v[3])), "r"(pack_bf16(v[4], v[5])), "r"(pack_bf16(v[6], v[7]))
                       : "memory");
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        ptx::tma_store_2d(&tmap_out, col0, m0 + ew * 32, slab);
        ptx::tma_store_commit();
      }
      __syncwarp();
      // the next tile's residual rows are requested later (after the next GELU's math): waiting here for the store
      // to have read the slab cost ~2000 cycles per tile (event trace)
      resid_tile = tile + num_pairs < num_tiles ? tile + num_pairs : -1;
    };
    // Order of work at a tile boundary (event trace: with the drain placed right after a warp's last chunk, the team
    // whose last chunk is not the tile's last idled ~3400 cycles waiting for O while S of the next tile sat ready,
    // and GEMM2 of the next tile started 6400 cycles after the last one of this tile):
    //   * a warp whose last chunk of the tile IS the tile's last drains next (O completes ~400 cycles later);
    //   * the other team first runs the GELU of its first chunk of the next tile (GEMM1 runs ahead across tiles, its S
    //     is ready; its H store waits for the last GEMM2 of this tile, not for any drain), then drains.
    uint32_t g = 0;
    int drain_next = 0, last_own_tl = -1, last_own_j = -1;
    for (int tl = 0; tl < my_tiles; ++tl) {
      for (int j = 0; j < nch; ++j, ++g) {
        const uint32_t b = g & 1u, use = (g >> 1) & 1u;
        if (b != team) continue;
        bool deferred = false;
        if (has_sub) {
          while (drain_next < tl - 1) drain(drain_next++);
          if (drain_next == tl - 1) {
            if (last_own_tl == tl - 1 && last_own_j == nch - 1) drain(drain_next++);
            else deferred = true;
          }
        }
        const int treg = warp_idx == 4 ? 2 : (warp_idx == 12 ? 3 : 0);
        if (treg > 0) ev(treg, 100 + static_cast<int>(g));
        ptx::mbar_wait(sfull(b), use);
        ptx::tc_fence_after();
        if (treg > 0) ev(treg, 200 + static_cast<int>(g));
        const uint32_t* b1h = s_b1h + (j * HC + half * 64) / 2;
        uint32_t hq[32];
#pragma unroll
        for (int hp = 0; hp < 2; ++hp) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(lane_base + (b ? TMEM_S1 : TMEM_S0) + half * 64u + static_cast<uint32_t>(hp * 32), r);
          ptx::tmem_ld_wait();
          if (hp == 1) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_cluster(b ? sempty_leader1 : sempty_leader0);  // S[b] is in registers
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint4 b4 = reinterpret_cast<const uint4*>(b1h + hp * 16)[q];
            hq[hp * 16 + 4 * q] = gelu_half_f16x2_b(__uint_as_float(r[8 * q]), __uint_as_float(r[8 * q + 1]), b4.x);
            hq[hp * 16 + 4 * q + 1] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 2]), __uint_as_float(r[8 * q + 3]), b4.y);
            hq[hp * 16 + 4 * q + 2] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 4]), __uint_as_float(r[8 * q + 5]), b4.z);
            hq[hp * 16 + 4 * q + 3] = gelu_half_f16x2_b(__uint_as_float(r[8 * q + 6]), __uint_as_float(r[8 * q + 7]), b4.w);
          }
        }
        if (treg > 0) ev(treg, 300 + static_cast<int>(g));
        if (resid_tile >= 0) {   // slab is free once the last store has read it: land the next residual rows
          if (lane == 0) request_resid(resid_tile);
          resid_tile = -1;
        }
        // GEMM2 of the previous chunk (the other team's) has finished reading this half of H
        if (g > 0) {
          ptx::mbar_wait(hempty(b ^ 1u, half), ((g - 1u) >> 1) & 1u);
          ptx::tc_fence_after();
        }
        if (treg > 0) ev(treg, 400 + static_cast<int>(g));
        ptx::tmem_st_32x32(lane_base + TMEM_H + half * 32u, hq);
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(hfull_leader);
        if (treg > 0) ev(treg, 500 + static_cast<int>(g));
        if (deferred) drain(drain_next++);
        last_own_tl = tl;
        last_own_j = j;
      }
    }
    if (has_sub)
      while (drain_next < my_tiles) drain(drain_next++);
    if (has_sub && lane == 0) ptx::tma_store_wait<0>();
  }

  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp_idx == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, 512);
  }
}

template <int C>
int launch_ffn_tmemh(const FfnFusedArgs& a, cudaStream_t stream) {
  using Cfg = FfnTCfg<C>;
  auto kfn = ffn_fused_tmemh_kernel<C>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), Cfg::SMEM_BYTES)) return rc;
  CUtensorMap tx, tw1, tw2, to, tr;
  if (int rc = make_tmap_bf16(&tx, a.x, a.M, C, C, FM)) return rc;
  if (int rc = make_tmap_bf16(&tw1, a.w1, a.hidden, C, C, HC / 2)) return rc;
  if (int rc = make_tmap_bf16(&tw2, a.w2, C, a.hidden, a.hidden, C / 2)) return rc;  // fp16 bits, same geometry
  if (int rc = make_tmap_bf16(&to, a.out, a.M, C, C, 32)) return rc;
  if (int rc = make_tmap_bf16(&tr, a.resid, a.M, C, C, 32)) return rc;
  FfnParams p;
  p.M = a.M; p.hidden = a.hidden; p.b1 = a.b1; p.b2 = a.b2;
  p.resid = static_cast<const __nv_bfloat16*>(a.resid); p.ldr = C;
  p.trace = nullptr;
#ifdef FVLA_FFN_TRACE_BUILD
  static int trace_calls = 0;
  const bool tr_now = ++trace_calls == 10;
  if (tr_now) {
    FVLA_CUDA_CHECK(cudaMalloc(&p.trace, 4 * 512 * 8));
    FVLA_CUDA_CHECK(cudaMemset(p.trace, 0, 4 * 512 * 8));
  }
#endif
  const int tiles = ceil_div(a.M, FPAIR_M);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(FFN_THREADS), Cfg::SMEM_BYTES, stream, tx, tw1, tw2, to, tr, p));
#ifdef FVLA_FFN_TRACE_BUILD
  if (tr_now) {   // build with -DFVLA_FFN_TRACE_BUILD: event timeline of launch #10 on stderr
    static unsigned long long h[4 * 512];
    FVLA_CUDA_CHECK(cudaStreamSynchronize(stream));
    FVLA_CUDA_CHECK(cudaMemcpy(h, p.trace, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(p.trace);
    const unsigned long long t0 = h[1];
    for (int r = 0; r < 4; ++r)
      for (int i = 0; i < 254 && h[r * 512 + 2 * i] != 0; ++i)
        fprintf(stderr, "TR %d %llu %lld\n", r, h[r * 512 + 2 * i], static_cast<long long>(h[r * 512 + 2 * i + 1] - t0));
  }
#endif
  return 0;
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
  p.trace = nullptr;
  const int tiles = ceil_div(a.M, FPAIR_M);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(FFN_THREADS), Cfg::SMEM_BYTES, stream, tx, tw1, tw2, to, p));
  return 0;
}

}  // namespace

bool ffn_fused_supported(int dtype, int C, int hidden) {
  if (ffn_wide_supported(dtype, C, hidden)) return true;
  return dtype == DT_BF16 && (C == 96 || C == 192) && hidden % HC == 0 && hidden >= HC && hidden <= 4 * C;
}

int ffn_fused(const FfnFusedArgs& a, cudaStream_t stream) {
  if (a.C == 384) return ffn_wide(a, stream);
  FVLA_REQUIRE(a.M > 0 && ffn_fused_supported(DT_BF16, a.C, a.hidden), "ffn_fused: unsupported shape");
  FVLA_REQUIRE(a.b1 != nullptr && a.b2 != nullptr && a.resid != nullptr, "ffn_fused: biases and residual required");
  static const bool smem_h = std::getenv("FVLA_FFN_SMEM_H") != nullptr;    // A/B switches for profiling
  static const bool teams = std::getenv("FVLA_FFN_NO_TEAMS") == nullptr;
  if (!smem_h) return a.C == 96 ? launch_ffn_tmemh<96>(a, stream) : launch_ffn_tmemh<192>(a, stream);
  if (a.C == 96) return teams ? launch_ffn<96, true>(a, stream) : launch_ffn<96, false>(a, stream);
  return teams ? launch_ffn<192, true>(a, stream) : launch_ffn<192, false>(a, stream);
}

}  // namespace fvla
