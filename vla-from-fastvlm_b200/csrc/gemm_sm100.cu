// bf16 GEMM for sm_100a:  D[M,N] = epilogue( A[M,K] * W[N,K]^T )
//
// This is the 1x1-conv / nn.Linear workhorse of the FastVLA forward (FastViTHD ConvFFN fc1/fc2,
// patch-embed pointwise convs, MHSA qkv/proj, mm_projector, Qwen2 q/k/v/o/gate/up/down).
// Activations are NHWC / token-major, so A is row-major [M, K]; torch keeps Linear and 1x1-conv
// weights as [N, K] — both operands are K-major and go to the tensor core untouched.
//
// Structure (one persistent CTA per SM, 640 threads):
//   warp 0     TMA producer: A/W tiles -> 128B-swizzled shared memory ring (mbarrier full/empty)
//   warp 1     MMA issuer: one thread issues tcgen05.mma (128 x BLOCK_N x 16), accumulators in TMEM,
//              two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warp 2     TMEM allocator
//   warps 4-19 epilogue (four warps per TMEM lane quarter, each taking a quarter of the columns):
//              tcgen05.ld -> row-scale / bias / activation / residual / SwiGLU -> bf16 ->
//              swizzled staging tile -> TMA store (bounds clipped by the tensor map)
//
//
// Edges need no special code: TMA zero-fills out-of-bounds loads (M, N and K tails) and clips stores.
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx_sm100.cuh"
#include "kernels.h"
#include "tma_host.h"
#include "epilogue_math.cuh"

#include <mutex>

namespace fvla {

namespace {

constexpr int BLOCK_M = 128;   // rows per CTA
constexpr int PAIR_M = 256;    // rows per CTA pair = UMMA M under cta_group::2
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int EPI_WPQ = 4;                        // epilogue warps per TMEM lane quarter
constexpr int EPI_THREADS = 4 * EPI_WPQ * 32;     // 512
constexpr int GEMM_THREADS = 128 + EPI_THREADS;   // 640
constexpr int SLAB_BYTES = 32 * 64 * 2;          // one warp's 32-row x 64-column staging slab
constexpr int STAGING_BYTES = 4 * EPI_WPQ * SLAB_BYTES;  // one private slab per epilogue warp (64 KB)

template <int BLOCK_N> struct GemmCfg {
  static constexpr int HALF_N = BLOCK_N / 2;           // rows of the W tile each CTA of the pair stages
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = HALF_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;  // per CTA
  static constexpr int STAGES = (BLOCK_N == 256) ? 5 : (BLOCK_N == 192 ? 5 : (BLOCK_N == 128 ? 6 : 8));
  static constexpr int TMEM_COLS = (2 * BLOCK_N <= 128) ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512);
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + BAR_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB dynamic shared memory limit");
};

struct EpiParams {
  int M, N, K;
  const float* bias;           // [N] or null
  const float* row_scale;      // [M] or null (applied to the accumulator before the bias)
  const __nv_bfloat16* resid;  // [M, ldr] or null (added after the activation)
  int ldr;
  int act;
  int ab_f16;                  // operands are fp16 (instruction descriptor formats), else bf16
  int out_f32;                 // D and resid are FP32 (the decoder's residual stream): two 32-column TMA stores per sub-tile
  const uint32_t* rope_tab;    // [rope_T][32] (cos, sin) fp16 pairs or null: RoPE on output columns < rope_cols
  int rope_T, rope_cols;
  int split_k;                 // >= 1
};

using namespace epi;

template <int BLOCK_N, bool SWIGLU, bool OUT_F32, bool ROPE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a,
                         const __grid_constant__ CUtensorMap tmap_w,
                         const __grid_constant__ CUtensorMap tmap_d, const EpiParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  // accumulator columns consumed per 64-column output sub-tile
  constexpr int ACC_PER_SUB = SWIGLU ? 128 : 64;
  static_assert(BLOCK_N % ACC_PER_SUB == 0, "BLOCK_N must cover whole output sub-tiles");
  constexpr int NUM_SUB = BLOCK_N / ACC_PER_SUB;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t smem_a0 = smem_base;
  const uint32_t smem_b0 = smem_base + STAGES * Cfg::A_BYTES;
  const uint32_t smem_stage0 = smem_base + STAGES * Cfg::STAGE_BYTES;
  const uint32_t bar_base = smem_stage0 + STAGING_BYTES;
  // barrier layout (8 B each): full[STAGES], empty[STAGES], tmem_full[2], tmem_empty[2], tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  auto tfull_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_ptr_smem = bar_base + 8u * (2 * STAGES + 4);

  const int warp_idx = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // A CTA pair (cluster of 2, one TPC) works on one 256 x BLOCK_N tile: CTA r stages rows [128r, 128r+128) of A
  // and rows [r*BLOCK_N/2, (r+1)*BLOCK_N/2) of the W tile; the leader issues cta_group::2 MMAs that read both
  // shared memories and write each CTA's 128 accumulator rows into that CTA's own TMEM.
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
  const int num_m = (p.M + PAIR_M - 1) / PAIR_M;
  const int num_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int num_kb_all = (p.K + BLOCK_K - 1) / BLOCK_K;
  // split-K: a work item is (tile, split); split s covers k-blocks [s * kb_per, min((s + 1) * kb_per, num_kb_all))
  const int kb_per = (num_kb_all + p.split_k - 1) / p.split_k;
  const int num_tiles = num_m * num_n * p.split_k;

  if (warp_idx == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_d);
  }
  if (warp_idx == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(full_bar(s), 2);    // leader's copy is the live one: one arrive.expect_tx per CTA
      ptx::mbar_init(empty_bar(s), 1);   // multicast tcgen05.commit from the leader
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(tfull_bar(a), 1);                 // multicast tcgen05.commit from the leader
      ptx::mbar_init(tempty_bar(a), 2 * 4 * EPI_WPQ);  // leader's copy: one arrive per epilogue warp of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp_idx == 2) {
    ptx::tmem_alloc_pair(tmem_ptr_smem, Cfg::TMEM_COLS);
    ptx::tmem_relinquish_pair();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  // Still under the previous kernel's tail: pull this CTA's half of the W panels of its FIRST work item into L2.  The
  // weights are constants, so this needs no ordering against the previous grid; at small M (the b = 1 prefill) they
  // are the whole HBM traffic of the launch and their DRAM round trip otherwise starts only after the wait.
  if (warp_idx == 0 && lane == 0 && pair_id < num_tiles) {
    const int tile = pair_id / p.split_k, ks = pair_id - tile * p.split_k;
    const int n0 = (tile % num_n) * BLOCK_N + static_cast<int>(cta_rank) * Cfg::HALF_N;
    const int kb_end = min(num_kb_all, (ks + 1) * kb_per);
    for (int kb = ks * kb_per; kb < kb_end; ++kb) ptx::tma_prefetch_2d(&tmap_w, kb * BLOCK_K, n0);
  }
  pdl_sync();  // everything above is input-independent and overlaps the previous kernel's tail

  if (warp_idx == 0) {
    // ===================== TMA producer (both CTAs) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = pair_id; work < num_tiles; work += num_pairs) {
        const int tile = work / p.split_k, ks = work - tile * p.split_k;
        const int m0 = (tile / num_n) * PAIR_M + static_cast<int>(cta_rank) * BLOCK_M;
        const int n0 = (tile % num_n) * BLOCK_N + static_cast<int>(cta_rank) * Cfg::HALF_N;
        const int kb_end = min(num_kb_all, (ks + 1) * kb_per);
        for (int kb = ks * kb_per; kb < kb_end; ++kb) {
          ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
          const uint32_t full_leader = ptx::mapa_rank(full_bar(stage), 0);
          ptx::mbar_arrive_expect_tx_cluster(full_leader, Cfg::STAGE_BYTES);
          ptx::tma_load_2d_pair(smem_a0 + stage * Cfg::A_BYTES, &tmap_a, kb * BLOCK_K, m0, full_leader);
          ptx::tma_load_2d_pair(smem_b0 + stage * Cfg::B_BYTES, &tmap_w, kb * BLOCK_K, n0, full_leader);
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp_idx == 1) {
    // ===================== MMA issuer (leader CTA only) =====================
    // warp-uniform control flow, tcgen05 instructions under elect_one: descriptors stay in uniform registers and
    // the four UTCHMMAs of a k-block issue back to back
    if (cta_rank == 0) {
      const uint32_t idesc = p.ab_f16 ? ptx::make_idesc_f16(PAIR_M, BLOCK_N) : ptx::make_idesc_bf16(PAIR_M, BLOCK_N);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int work = pair_id; work < num_tiles; work += num_pairs) {
        const int ks = work % p.split_k;
        const int num_kb = min(num_kb_all, (ks + 1) * kb_per) - ks * kb_per;
        ptx::mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BLOCK_N);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full_bar(stage), phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_kmajor_sw128_desc(smem_a0 + stage * Cfg::A_BYTES);
          const uint64_t db = ptx::make_kmajor_sw128_desc(smem_b0 + stage * Cfg::B_BYTES);
          if (ptx::elect_one()) {
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 16 bf16 = 32 B inside the swizzle row: +2 in the (addr >> 4) field
              ptx::umma_bf16_pair(tmem_d, da + static_cast<uint64_t>(2 * k),
                                  db + static_cast<uint64_t>(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            ptx::umma_commit_pair(empty_bar(stage), 3);  // frees this smem slot in both CTAs once the MMAs retire
            if (kb == num_kb - 1) ptx::umma_commit_pair(tfull_bar(acc), 3);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if (warp_idx >= 4) {
    // ===================== epilogue =====================
    // A warp may only read its own TMEM lane quarter (warp_idx % 4), which is also the scheduler it issues
    // on: the four warps of a quarter share an issue port.  Each warp therefore owns WHOLE 32-row x 64-column
    // output sub-tiles (private 4 KB staging slab, private TMA store), the sub-tiles of a tile rotating over
    // the four warps of the quarter.  No barrier joins the warps, so they drift apart and one warp's TMEM
    // load / store-drain / residual latency is covered by the others' math (the earlier shared-slab version
    // kept all four in lock step: the latencies of every phase added up, see profiles/r01_gemm_epilogue_ab.txt).
    const int ew = warp_idx & 3;            // TMEM lane quarter
    const int grp = (warp_idx - 4) >> 2;    // which of the quarter's four warps
    const int row = ew * 32 + lane;         // tile row == TMEM lane
    const uint32_t slab = smem_stage0 + static_cast<uint32_t>(warp_idx - 4) * SLAB_BYTES;
    const uint32_t sbase = slab + lane * 128;  // this thread's row inside the slab
    int acc = 0;
    uint32_t acc_phase = 0;
    uint32_t sub_counter = 0;
    const int n_out_total = SWIGLU ? p.N / 2 : p.N;
    const uint32_t tempty_leader0 = ptx::mapa_rank(tempty_bar(0), 0);
    const uint32_t tempty_leader1 = ptx::mapa_rank(tempty_bar(1), 0);
    for (int work = pair_id; work < num_tiles; work += num_pairs) {
      const int tile = work / p.split_k;
      const bool first_split = work - tile * p.split_k == 0;  // the bias is added by exactly one split
      const int m0 = (tile / num_n) * PAIR_M + static_cast<int>(cta_rank) * BLOCK_M;
      const int n0 = (tile % num_n) * BLOCK_N;
      const int m = m0 + row;
      float rs = 1.0f;
      if (p.row_scale != nullptr && m < p.M) rs = __ldg(p.row_scale + m);

      // This warp's sub-tile of the tile (the rotation gives every warp at most one): known before the
      // accumulator is ready, so the slab hand-back and the residual loads are issued under the wait for the MMAs.
      const int sub = static_cast<int>((static_cast<uint32_t>(grp) - sub_counter) & 3u);
      sub_counter += NUM_SUB;
      const int acc_col0 = sub * ACC_PER_SUB;                    // first accumulator column
      const int out_col0 = (SWIGLU ? (n0 / 2) : n0) + sub * 64;  // first output column
      const bool has_sub = sub < NUM_SUB && out_col0 < n_out_total;
      if (has_sub) {
        // the TMA store this warp issued from its slab last time must have finished reading it
        if (lane == 0) ptx::tma_store_wait_read<0>();
        __syncwarp();
        if (ROPE && out_col0 < p.rope_cols) {
          // RoPE head: the (cos, sin) rows of this warp's 32 positions (32 x 128 B of fp16 pairs) go through the slab,
          // loaded coalesced under the wait for the MMAs; each thread reads its own row back below
          const int base_pos = (m0 + ew * 32) % p.rope_T;
          uint4 tt[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 32 + lane;
            const int rrow = idx >> 3, rchunk = idx & 7;
            int pos = base_pos + rrow;
            if (pos >= p.rope_T) pos -= p.rope_T;
            tt[i] = __ldg(reinterpret_cast<const uint4*>(p.rope_tab + static_cast<size_t>(pos) * 32) + rchunk);
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 32 + lane;
            const int rrow = idx >> 3, rchunk = idx & 7;
            const uint32_t dst = slab + rrow * 128 + ((rchunk ^ (rrow & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(tt[i].x), "r"(tt[i].y),
                         "r"(tt[i].z), "r"(tt[i].w)
                         : "memory");
          }
          __syncwarp();
        }
        if (!SWIGLU && !OUT_F32 && p.resid != nullptr) {
          // residual sub-tile (32 rows x 128 B): coalesced 16-byte loads into the slab
          uint4 rr[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 32 + lane;
            const int rrow = idx >> 3, rchunk = idx & 7;
            const int gm = m0 + ew * 32 + rrow, gc = out_col0 + rchunk * 8;
            rr[i] = make_uint4(0u, 0u, 0u, 0u);
            if (gm < p.M && gc + 8 <= p.N)
              rr[i] = __ldg(reinterpret_cast<const uint4*>(p.resid + static_cast<size_t>(gm) * p.ldr + gc));
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int idx = i * 32 + lane;
            const int rrow = idx >> 3, rchunk = idx & 7;
            const uint32_t dst = slab + rrow * 128 + ((rchunk ^ (rrow & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(rr[i].x), "r"(rr[i].y),
                         "r"(rr[i].z), "r"(rr[i].w)
                         : "memory");
          }
          __syncwarp();
        }
      }

      ptx::mbar_wait(tfull_bar(acc), acc_phase);
      ptx::tc_fence_after();
      const uint32_t tmem_acc =
          tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + static_cast<uint32_t>(acc * BLOCK_N);

      if constexpr (ROPE) {
        if (has_sub && out_col0 < p.rope_cols) {
          // ---- qkv projection epilogue with RoPE (Qwen2 apply_rotary_pos_emb, rotate_half): this sub-tile is one
          // head of 64 columns and column j pairs with j + 32, so the halves leave TMEM 16 columns at a time ----
          uint4 hold0 = make_uint4(0u, 0u, 0u, 0u), hold1 = hold0;
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            uint32_t ra[16], rb[16];
            ptx::tmem_ld_32x16(tmem_acc + static_cast<uint32_t>(acc_col0 + 16 * q), ra);
            ptx::tmem_ld_32x16(tmem_acc + static_cast<uint32_t>(acc_col0 + 32 + 16 * q), rb);
            uint32_t cs[16];  // this row's (cos, sin) pairs for rotation indices 16q .. 16q+15
#pragma unroll
            for (int c = 0; c < 4; ++c)
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                           : "=r"(cs[4 * c]), "=r"(cs[4 * c + 1]), "=r"(cs[4 * c + 2]), "=r"(cs[4 * c + 3])
                           : "r"(sbase + (((4 * q + c) ^ (lane & 7)) << 4)));
            const int nb = n0 + acc_col0 + 16 * q;
            float4 ba[4], bb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              ba[j] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
              bb[j] = p.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(p.bias + nb + 32) + j) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            ptx::tmem_ld_wait();
            float lo[16], hi[16];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a4[4] = {ba[j].x, ba[j].y, ba[j].z, ba[j].w}, b4[4] = {bb[j].x, bb[j].y, bb[j].z, bb[j].w};
#pragma unroll
              for (int t = 0; t < 4; ++t) {
                const __half2 h2 = *reinterpret_cast<const __half2*>(&cs[4 * j + t]);
                const float c = __low2float(h2), sn = __high2float(h2);
                const float x1 = __uint_as_float(ra[4 * j + t]) + a4[t];
                const float x2 = __uint_as_float(rb[4 * j + t]) + b4[t];
                lo[4 * j + t] = x1 * c - x2 * sn;
                hi[4 * j + t] = fmaf(x2, c, x1 * sn);
              }
            }
            // Outputs overwrite this thread's own table row: q == 0 consumed table chunks 0-3 and produces output
            // chunks 0,1 (low half) and 4,5 (high half) — 4,5 still hold table pairs for q == 1, so they wait in registers
            if (q == 0) {
#pragma unroll
              for (int c = 0; c < 2; ++c)
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (((c) ^ (lane & 7)) << 4)),
                             "r"(pack_bf16(lo[8 * c], lo[8 * c + 1])), "r"(pack_bf16(lo[8 * c + 2], lo[8 * c + 3])),
                             "r"(pack_bf16(lo[8 * c + 4], lo[8 * c + 5])), "r"(pack_bf16(lo[8 * c + 6], lo[8 * c + 7]))
                             : "memory");
              hold0 = make_uint4(pack_bf16(hi[0], hi[1]), pack_bf16(hi[2], hi[3]), pack_bf16(hi[4], hi[5]), pack_bf16(hi[6], hi[7]));
              hold1 = make_uint4(pack_bf16(hi[8], hi[9]), pack_bf16(hi[10], hi[11]), pack_bf16(hi[12], hi[13]), pack_bf16(hi[14], hi[15]));
            } else {
#pragma unroll
              for (int c = 0; c < 2; ++c) {
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (((2 + c) ^ (lane & 7)) << 4)),
                             "r"(pack_bf16(lo[8 * c], lo[8 * c + 1])), "r"(pack_bf16(lo[8 * c + 2], lo[8 * c + 3])),
                             "r"(pack_bf16(lo[8 * c + 4], lo[8 * c + 5])), "r"(pack_bf16(lo[8 * c + 6], lo[8 * c + 7]))
                             : "memory");
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (((6 + c) ^ (lane & 7)) << 4)),
                             "r"(pack_bf16(hi[8 * c], hi[8 * c + 1])), "r"(pack_bf16(hi[8 * c + 2], hi[8 * c + 3])),
                             "r"(pack_bf16(hi[8 * c + 4], hi[8 * c + 5])), "r"(pack_bf16(hi[8 * c + 6], hi[8 * c + 7]))
                             : "memory");
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (((4) ^ (lane & 7)) << 4)),
                           "r"(hold0.x), "r"(hold0.y), "r"(hold0.z), "r"(hold0.w) : "memory");
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sbase + (((5) ^ (lane & 7)) << 4)),
                           "r"(hold1.x), "r"(hold1.y), "r"(hold1.z), "r"(hold1.w) : "memory");
            }
          }
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_d, out_col0, m0 + ew * 32, slab);
            ptx::tma_store_commit();
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
          if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
          continue;
        }
      }
      if (has_sub) {
#pragma unroll
        for (int hp = 0; hp < ACC_PER_SUB / 32; ++hp) {
          uint32_t r[32];
          ptx::tmem_ld_32x32(tmem_acc + static_cast<uint32_t>(acc_col0 + hp * 32), r);
          if constexpr (OUT_F32) {
            // FP32 output: the 4 KB slab holds 32 rows x 32 columns, so a sub-tile leaves in two halves
            if (hp > 0) {
              if (lane == 0) ptx::tma_store_wait_read<0>();
              __syncwarp();
            }
          }
          // bias for these 32 columns: issued under the TMEM load so the two latencies overlap
          const int nb = n0 + acc_col0 + hp * 32;  // global column of v[0]
          const bool use_bias = p.bias != nullptr && first_split;
          const bool bias_vec = !SWIGLU && use_bias && nb + 32 <= p.N;
          float4 bq[8];
          if (bias_vec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) bq[j] = __ldg(reinterpret_cast<const float4*>(p.bias + nb) + j);
          }
          ptx::tmem_ld_wait();
          float v[32];
          if (p.row_scale != nullptr) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) * rs;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          }
          if constexpr (SWIGLU) {
            // columns are (gate, up) pairs: 32 accumulators -> 16 outputs = two 16-byte chunks
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              uint32_t o[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float a0 = silu_fast(v[16 * c + 4 * j]) * v[16 * c + 4 * j + 1];
                const float a1 = silu_fast(v[16 * c + 4 * j + 2]) * v[16 * c + 4 * j + 3];
                o[j] = pack_bf16(a0, a1);
              }
              const int chunk = 2 * hp + c;
              const uint32_t dst = sbase + ((chunk ^ (lane & 7)) << 4);
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(o[0]), "r"(o[1]),
                           "r"(o[2]), "r"(o[3])
                           : "memory");
            }
          } else {
            if (use_bias) {
              if (bias_vec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[4 * j] += bq[j].x; v[4 * j + 1] += bq[j].y; v[4 * j + 2] += bq[j].z; v[4 * j + 3] += bq[j].w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (nb + j < p.N) v[j] += __ldg(p.bias + nb + j);
              }
            }
            if (p.act == ACT_GELU_HALF_F16) {
              // fp16 hidden tensor: the whole activation on packed half pairs (epilogue_math.cuh), no residual
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                const int chunk = hp * 4 + c;
                const uint32_t dst = sbase + ((chunk ^ (lane & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                             "r"(gelu_half_f16x2(v[8 * c], v[8 * c + 1])),
                             "r"(gelu_half_f16x2(v[8 * c + 2], v[8 * c + 3])),
                             "r"(gelu_half_f16x2(v[8 * c + 4], v[8 * c + 5])),
                             "r"(gelu_half_f16x2(v[8 * c + 6], v[8 * c + 7]))
                             : "memory");
              }
              continue;
            }
            if (p.act == ACT_GELU_HALF) {
              gelu_half_hybrid(v);
            } else if (p.act == ACT_GELU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] *= 0.5f;
              gelu_half_hybrid(v);
            } else if (p.act == ACT_SILU) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = silu_fast(v[j]);
            }
            if constexpr (OUT_F32) {
              // The residual stream is updated IN PLACE (resid == D, checked on the host): the slab carries the
              // fp32 accumulators and the TMA store adds them to D in L2 (cp.reduce.async.bulk .add) — no
              // residual read in the epilogue at all.
#pragma unroll
              for (int c = 0; c < 8; ++c) {
                const uint32_t dst = sbase + ((c ^ (lane & 7)) << 4);
                asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(v[4 * c]), "f"(v[4 * c + 1]),
                             "f"(v[4 * c + 2]), "f"(v[4 * c + 3])
                             : "memory");
              }
              ptx::fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (p.resid != nullptr) ptx::tma_reduce_add_2d(&tmap_d, out_col0 + hp * 32, m0 + ew * 32, slab);
                else ptx::tma_store_2d(&tmap_d, out_col0 + hp * 32, m0 + ew * 32, slab);
                ptx::tma_store_commit();
              }
              continue;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const int chunk = hp * 4 + c;
              const uint32_t dst = sbase + ((chunk ^ (lane & 7)) << 4);
              if (p.resid != nullptr) {  // residual already sits at this slot (own row: no cross-lane hazard)
                uint32_t w0, w1, w2, w3;
                asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                             : "r"(dst));
                const uint32_t w[4] = {w0, w1, w2, w3};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                  v[8 * c + 2 * t] += __uint_as_float(w[t] << 16);
                  v[8 * c + 2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
                }
              }
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(dst),
                           "r"(pack_bf16(v[8 * c], v[8 * c + 1])),
                           "r"(pack_bf16(v[8 * c + 2], v[8 * c + 3])),
                           "r"(pack_bf16(v[8 * c + 4], v[8 * c + 5])),
                           "r"(pack_bf16(v[8 * c + 6], v[8 * c + 7]))
                           : "memory");
            }
          }
        }
        if constexpr (!OUT_F32) {
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_2d(&tmap_d, out_col0, m0 + ew * 32, slab);
            ptx::tma_store_commit();
          }
        }
      }
      // all TMEM reads of this accumulator stage are complete (tcgen05.wait::ld above)
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(acc == 0 ? tempty_leader0 : tempty_leader1);
      if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
    }
    if (lane == 0) ptx::tma_store_wait<0>();
  }

  // neither CTA may exit (or free TMEM) while the peer can still signal its barriers or the leader's MMAs read
  // its shared memory
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp_idx == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---- host side: tensor maps + launch --------------------------------------------------------

}  // namespace

TmaEncodeTiledFn tma_encode_fn() {
  static TmaEncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess) {
      fn = reinterpret_cast<TmaEncodeTiledFn>(sym);
    }
  });
  return fn;
}

namespace {

}  // namespace

// 2-D bf16 tensor [rows, cols] with row pitch ld (elements); box = box_rows x 64 cols, 128B swizzle.
int make_tmap_bf16(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld,
                   int box_rows) {
  TmaEncodeTiledFn fn = tma_encode_fn();
  FVLA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FVLA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA base must be 16-byte aligned");
  FVLA_REQUIRE((ld * 2) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {64u, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

namespace {

// FP32 output [rows, cols]: 32-row x 32-column (128-byte) store boxes through the same swizzled 4 KB slabs
int make_tmap_f32_store(CUtensorMap* out, void* ptr, long long rows, long long cols, long long ld) {
  TmaEncodeTiledFn fn = tma_encode_fn();
  FVLA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FVLA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0, "TMA base must be 16-byte aligned");
  FVLA_REQUIRE((ld * 4) % 16 == 0, "TMA row pitch must be a multiple of 16 bytes");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  cuuint32_t box[2] = {32u, 32u};
  cuuint32_t estr[2] = {1u, 1u};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 D) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

template <int BLOCK_N, bool SWIGLU, bool OUT_F32 = false, bool ROPE = false>
int launch_gemm(const GemmArgs& g, cudaStream_t stream) {
  static_assert(!ROPE || (!SWIGLU && !OUT_F32), "the RoPE epilogue is a plain bf16 epilogue");
  using Cfg = GemmCfg<BLOCK_N>;
  auto kfn = gemm_bf16_tcgen05_kernel<BLOCK_N, SWIGLU, OUT_F32, ROPE>;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), Cfg::SMEM_BYTES)) return rc;
  CUtensorMap ta, tw, td;
  if (int rc = make_tmap_bf16(&ta, g.A, g.M, g.K, g.lda, BLOCK_M)) return rc;
  if (int rc = make_tmap_bf16(&tw, g.W, g.N, g.K, g.ldw, Cfg::HALF_N)) return rc;
  const int n_out = SWIGLU ? g.N / 2 : g.N;
  if (OUT_F32) {
    if (int rc = make_tmap_f32_store(&td, g.D, g.M, n_out, g.ldd)) return rc;
  } else if (int rc = make_tmap_bf16(&td, g.D, g.M, n_out, g.ldd, 32)) {  // per-quarter 32-row stores
    return rc;
  }
  EpiParams ep;
  ep.M = g.M; ep.N = g.N; ep.K = g.K;
  ep.bias = g.bias; ep.row_scale = g.row_scale;
  ep.resid = static_cast<const __nv_bfloat16*>(g.resid); ep.ldr = g.ldr; ep.act = g.act; ep.ab_f16 = g.ab_f16; ep.out_f32 = g.out_f32;
  ep.rope_tab = g.rope_tab; ep.rope_T = g.rope_T; ep.rope_cols = g.rope_cols;
  const int tiles = ceil_div(g.M, PAIR_M) * ceil_div(g.N, BLOCK_N);
  const int pairs = num_sms() / 2;
  // split-K (reduce-add epilogue only): every split must own at least one k-block
  int split = 1;
  if (OUT_F32 && g.split_k > 1 && g.resid != nullptr) {
    const int num_kb = ceil_div(g.K, BLOCK_K);
    const int kb_per = ceil_div(num_kb, g.split_k < num_kb ? g.split_k : num_kb);
    split = ceil_div(num_kb, kb_per);
  }
  ep.split_k = split;
  const long long work = static_cast<long long>(tiles) * split;
  const int grid = 2 * static_cast<int>(work < pairs ? work : pairs);
  FVLA_CUDA_CHECK(launch_pdl(kfn, dim3(grid), dim3(GEMM_THREADS), Cfg::SMEM_BYTES, stream, ta, tw, td, ep));
  return 0;
}

// Pick the N tile by modelled time: waves over the 74 CTA pairs x time per tile.  Per-tile time ~ BLOCK_N / eff
// (narrow tiles move more operand bytes per flop through the same smem ring: BLOCK_N = 128 sustains ~0.78 of the
// 256-wide rate on a long-K GEMM, 64 about half) plus a fixed pipeline fill / epilogue term.  For large M this
// reduces to "least padded work per unit efficiency" (N = 896 -> 4 x 256, not 7 x 128); for small M (the b=1
// select_action prefill, M = 272) it prefers narrow tiles that spread the N dimension over more SMs — a 2 x 4 grid
// of 256-wide tiles leaves 66 of 74 pairs idle while each busy pair walks the whole K loop.
int pick_block_n(int M, int N, bool swiglu) {
  const int cands_plain[4] = {256, 192, 128, 64};
  const double eff_plain[4] = {1.0, 0.96, 0.78, 0.5};
  const int cands_glu[2] = {256, 128};
  const double eff_glu[2] = {1.0, 0.78};
  const int* cands = swiglu ? cands_glu : cands_plain;
  const double* eff = swiglu ? eff_glu : eff_plain;
  const int nc = swiglu ? 2 : 4;
  const int pairs = num_sms() / 2;
  const long long m_tiles = ceil_div(M, PAIR_M);
  int best = cands[0];
  double best_cost = -1.0;
  for (int i = 0; i < nc; ++i) {
    const int bn = cands[i];
    const long long tiles = m_tiles * ceil_div(N, bn);
    const double waves = static_cast<double>((tiles + pairs - 1) / pairs);
    const double cost = waves * (bn / eff[i] + 48.0);
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

}  // namespace

int gemm_bf16(const GemmArgs& g, cudaStream_t stream) {
  FVLA_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "empty GEMM");
  FVLA_REQUIRE(g.K % 8 == 0 && g.lda % 8 == 0 && g.ldw % 8 == 0 && g.ldd % 8 == 0,
               "bf16 GEMM needs K and leading dimensions in multiples of 8 elements");
  if (g.swiglu) {
    FVLA_REQUIRE(g.N % 16 == 0 && g.bias == nullptr && g.resid == nullptr && g.act == ACT_NONE,
                 "SwiGLU epilogue takes interleaved gate/up columns and no bias/residual/activation");
  }
  if (g.resid != nullptr) FVLA_REQUIRE(g.ldr % 8 == 0, "residual pitch must be a multiple of 8");
  if (g.out_f32)
    FVLA_REQUIRE(!g.swiglu && g.act != ACT_GELU_HALF_F16 && (g.resid == nullptr || (g.resid == g.D && g.ldr == g.ldd)),
                 "fp32 output: plain / bias / activation epilogues; a residual must be D itself (in-place stream update)");
  if (g.act == ACT_GELU_HALF_F16)
    FVLA_REQUIRE(g.resid == nullptr && !g.swiglu, "the fp16-output GELU epilogue takes no residual");
  if (g.rope_tab != nullptr)
    FVLA_REQUIRE(g.rope_T >= 32 && g.rope_cols % 64 == 0 && g.rope_cols <= g.N && g.N % 64 == 0 && !g.swiglu &&
                     !g.out_f32 && g.act == ACT_NONE && g.resid == nullptr && g.row_scale == nullptr,
                 "RoPE epilogue: plain bf16 projection whose rotated columns are whole 64-wide heads, >= 32 positions");
  if (g.split_k > 1)
    FVLA_REQUIRE(g.out_f32 && g.resid == g.D && g.act == ACT_NONE,
                 "split-K needs the reduce-add epilogue (fp32 D updated in place)");
  const int bn = g.block_n > 0 ? g.block_n : pick_block_n(g.M, g.N, g.swiglu != 0);
  if (g.swiglu) {
    switch (bn) {
      case 256: return launch_gemm<256, true>(g, stream);
      case 128: return launch_gemm<128, true>(g, stream);
      default: break;
    }
  } else if (g.out_f32) {
    switch (bn) {
      case 256: return launch_gemm<256, false, true>(g, stream);
      case 192: return launch_gemm<192, false, true>(g, stream);
      case 128: return launch_gemm<128, false, true>(g, stream);
      case 64: return launch_gemm<64, false, true>(g, stream);
      default: break;
    }
  } else if (g.rope_tab != nullptr) {
    switch (bn) {
      case 256: return launch_gemm<256, false, false, true>(g, stream);
      case 192: return launch_gemm<192, false, false, true>(g, stream);
      case 128: return launch_gemm<128, false, false, true>(g, stream);
      case 64: return launch_gemm<64, false, false, true>(g, stream);
      default: break;
    }
  } else {
    switch (bn) {
      case 256: return launch_gemm<256, false>(g, stream);
      case 192: return launch_gemm<192, false>(g, stream);
      case 128: return launch_gemm<128, false>(g, stream);
      case 64: return launch_gemm<64, false>(g, stream);
      default: break;
    }
  }
  set_error("unsupported GEMM N tile " + std::to_string(bn));
  return 2;
}

}  // namespace fvla
