// FastViTHD MHSA on the 5th-generation tensor cores (sm_100a): non-causal attention, head_dim 32, N % 128 == 0
// (stage 3: N = 1024, 24 heads; stage 4: N = 256, 48 heads)  [EXT mci.py MHSA.forward; SURVEY App. A].
//
//   CTA = 128 queries of one (sample, head).  6 warps:
//     warps 0-3  softmax: thread = one query row = one TMEM lane.  Scores come out of TMEM with tcgen05.ld, the
//                probabilities go back INTO TMEM as bf16 pairs (tcgen05.st) where the next MMA reads them as its A
//                operand: P never touches shared memory or registers of another thread, no shuffles, no smem round trip
//     warp 4     TMA producer: Q once, then K/V tiles of 64 keys through a 3-stage mbarrier ring; the tiles are column
//                slices of the fused qkv buffer, 64-byte rows under the 64-byte swizzle
//     warp 5     MMA issuer (one elected thread):
//                  S[128x64]   = Q . K^T           tcgen05.mma SS, both operands K-major, K = 32 (two steps)
//                  O[128x32]  += P . V             tcgen05.mma TS: A = P from TMEM, B = V tile as it sits in memory
//                                                  ([key][dim] = MN-major B operand), K = 64 keys (four steps)
//                  L[128x16]  += P . 1             the softmax denominator on the tensor core too (a constant tile of
//                                                  ones): numerator and denominator see the same bf16-rounded P
//   TMEM (128 columns per CTA, so 3-4 CTAs share an SM and overlap each other's MMA / softmax / load phases):
//     [0,64) S (fp32), overwritten in place by P (bf16 pairs, columns [0,32)); [64,96) O; [96,112) L.
//
// Online softmax with a LAZY running maximum: the reference point m of a row only moves when the tile's maximum
// exceeds it by more than 2^8 — probabilities then stay <= 256, exact in fp32 / bf16 — so the O/L rescale
// (a TMEM read-modify-write) runs on the first tile and almost never again.  exp2 with the softmax scale folded into
// one FFMA2 per score pair.
//
// What bounds it: one MUFU.EX2 per score (16 / clk / SM); the MMAs (4 * N^2 * 32 flops per head) are ~4x cheaper
// than that at the tensor pipe's rate, which is why the mma.sync kernel this replaces (attention_v2.cu: scores, P and
// the softmax bookkeeping all in registers, ~8 issue slots per score) was issue-bound, not tensor-bound.
#include "common.cuh"
#include "epilogue_math.cuh"
#include "kernels.h"
#include "ptx_sm100.cuh"
#include "tma_host.h"

#include <cstdlib>

namespace fvla {
namespace {
using namespace epi;

constexpr int AQ = 128;          // queries per CTA
constexpr int AKV = 64;          // keys per tile
constexpr int AHD = 32;          // head_dim
constexpr int A_STAGES = 3;
constexpr int A_THREADS = 192;
constexpr int Q_BYTES = AQ * AHD * 2;       // 8 KB
constexpr int KV_BYTES = AKV * AHD * 2;     // 4 KB each for K and V
constexpr int ONES_BYTES = 1024;
constexpr int A_SMEM = Q_BYTES + A_STAGES * 2 * KV_BYTES + ONES_BYTES + 256 + 1024;
constexpr uint32_t T_S = 0, T_O = 64, T_L = 96, T_COLS = 128;
constexpr float LAZY_LOG2 = 8.0f;
#ifndef FVLA_ATTN_MIN_CTAS
#define FVLA_ATTN_MIN_CTAS 4
#endif
constexpr int A_MIN_CTAS = FVLA_ATTN_MIN_CTAS;  // 4 x 128 TMEM columns = the whole tensor memory of an SM

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(A_THREADS, A_MIN_CTAS)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
               const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int ld_o, int N,
               int kv_group, float scale_log2) {
  extern __shared__ uint8_t smem_attn[];
  const uint32_t base = (ptx::smem_u32(smem_attn) + 1023u) & ~1023u;
  const uint32_t s_q = base;
  const uint32_t s_kv = s_q + Q_BYTES;                        // [stage][K | V]
  const uint32_t s_ones = s_kv + A_STAGES * 2 * KV_BYTES;
  const uint32_t bars = s_ones + ONES_BYTES;
  const uint32_t q_full = bars;
  auto kv_full = [&](int s) { return bars + 8u + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 8u + 8u * (A_STAGES + s); };
  const uint32_t s_full = bars + 8u + 8u * (2 * A_STAGES);
  const uint32_t p_full = s_full + 8u, o_full = s_full + 16u;
  const uint32_t tmem_ptr_smem = s_full + 24u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const int row_base = b * N;
  const int nkv = N / AKV;

  if (warp == 5 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < A_STAGES; ++s) {
      ptx::mbar_init(kv_full(s), 1);
      ptx::mbar_init(kv_empty(s), 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 4);  // one arrival per softmax warp
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  // the constant B operand of the row-sum MMA: bf16 1.0 everywhere (any layout of all-ones is all-ones)
  for (int i = threadIdx.x; i < ONES_BYTES / 4; i += A_THREADS)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(s_ones + 4u * i), "r"(0x3F803F80u) : "memory");
  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_k);
      ptx::prefetch_tmap(&tmap_v);
    }
    ptx::tmem_alloc(tmem_ptr_smem, T_COLS);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();  // the generic-proxy stores of the ones tile, before the tensor core reads it
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  pdl_sync();

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, Q_BYTES);
      ptx::tma_load_2d(s_q, &tmap_q, h * AHD, row_base + q0, q_full);
      ptx::tma_load_2d(s_q + Q_BYTES / 2, &tmap_q, h * AHD, row_base + q0 + 64, q_full);
      const int hk = h / kv_group;
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        ptx::mbar_wait(kv_empty(stage), phase ^ 1u);
        ptx::mbar_arrive_expect_tx(kv_full(stage), 2 * KV_BYTES);
        ptx::tma_load_2d(s_kv + stage * 2 * KV_BYTES, &tmap_k, hk * AHD, row_base + j * AKV, kv_full(stage));
        ptx::tma_load_2d(s_kv + stage * 2 * KV_BYTES + KV_BYTES, &tmap_v, hk * AHD, row_base + j * AKV, kv_full(stage));
        if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(AQ, AKV);
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(AQ, AHD);
    constexpr uint32_t idesc_l = ptx::make_idesc_bf16(AQ, 16);
    const uint64_t dq = ptx::make_sw64_desc(s_q);
    const uint64_t d1 = ptx::make_sw64_desc(s_ones);
    int stage = 0;
    uint32_t phase = 0;
    auto issue_qk = [&](int st) {
      const uint64_t dk = ptx::make_sw64_desc(s_kv + st * 2 * KV_BYTES);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < AHD / 16; ++k)  // +32 B per K step inside the 64-byte swizzle row
          ptx::umma_bf16(tmem_base + T_S, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k),
                         idesc_qk, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(kv_full(0), 0);
    ptx::tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < nkv; ++j) {
      ptx::mbar_wait(p_full, j & 1);
      ptx::tc_fence_after();
      const uint64_t dv = ptx::make_sw64_desc(s_kv + stage * 2 * KV_BYTES + KV_BYTES);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < AKV / 16; ++k) {
          const uint32_t acc = (j > 0 || k > 0) ? 1u : 0u;
          // A: 16 keys of P = 8 TMEM columns; B: 16 key rows of V = 1024 B further
          ptx::umma_bf16_ts(tmem_base + T_O, tmem_base + T_S + 8u * k, dv + static_cast<uint64_t>(64 * k), idesc_pv, acc);
          ptx::umma_bf16_ts(tmem_base + T_L, tmem_base + T_S + 8u * k, d1, idesc_l, acc);
        }
        ptx::umma_commit(kv_empty(stage));  // K and V of this stage have been consumed once these MMAs retire
      }
      __syncwarp();
      if (++stage == A_STAGES) { stage = 0; phase ^= 1u; }
      if (j + 1 < nkv) {
        ptx::mbar_wait(kv_full(stage), phase);
        ptx::tc_fence_after();
        issue_qk(stage);  // overwrites S/P: ordered behind the P.V MMAs above in the tensor pipe
      } else {
        if (ptx::elect_one()) ptx::umma_commit(o_full);
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax (warps 0-3: TMEM lane quarter == warp index) =====================
    const uint32_t tl = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const f32x2 c2 = splat2(scale_log2);
    float mx = -INFINITY;  // reference point of this row's exponentials, in scaled log2 units
    for (int j = 0; j < nkv; ++j) {
      ptx::mbar_wait(s_full, j & 1);
      ptx::tc_fence_after();
      uint32_t s0[32], s1[32];
      ptx::tmem_ld_32x32(tl + T_S, s0);
      ptx::tmem_ld_32x32(tl + T_S + 32u, s1);
      ptx::tmem_ld_wait();
      float t0 = __uint_as_float(s0[0]), t1 = __uint_as_float(s0[1]), t2 = __uint_as_float(s1[0]),
            t3 = __uint_as_float(s1[1]);
#pragma unroll
      for (int i = 2; i < 32; i += 2) {
        t0 = fmaxf(t0, __uint_as_float(s0[i]));
        t1 = fmaxf(t1, __uint_as_float(s0[i + 1]));
        t2 = fmaxf(t2, __uint_as_float(s1[i]));
        t3 = fmaxf(t3, __uint_as_float(s1[i + 1]));
      }
      const float tm = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)) * scale_log2;
      const bool need = tm > mx + LAZY_LOG2;  // always on the first tile (mx = -inf)
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // move this row's reference point: O and L were accumulated relative to the old one
        // (rare: eight columns at a time keeps the 64 live scores + this path under the 4-CTAs-per-SM register budget)
        const float f = need ? ex2_approx(mx - tm) : 1.0f;
        uint32_t l;
        ptx::tmem_ld_32x1(tl + T_L, l);
#pragma unroll 1
        for (int c = 0; c < AHD; c += 8) {
          uint32_t o[8];
          ptx::tmem_ld_32x8(tl + T_O + c, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          ptx::tmem_st_32x8(tl + T_O + c, o);
        }
        ptx::tmem_st_32x1(tl + T_L, __float_as_uint(__uint_as_float(l) * f));
      }
      if (need) mx = tm;
      const f32x2 nm2 = splat2(-mx);
      uint32_t p[32];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a0, a1, b0, b1;
        upk2(fma2(pk2(__uint_as_float(s0[2 * i]), __uint_as_float(s0[2 * i + 1])), c2, nm2), a0, a1);
        upk2(fma2(pk2(__uint_as_float(s1[2 * i]), __uint_as_float(s1[2 * i + 1])), c2, nm2), b0, b1);
        p[i] = pack_bf16(ex2_approx(a0), ex2_approx(a1));
        p[16 + i] = pack_bf16(ex2_approx(b0), ex2_approx(b1));
      }
      ptx::tmem_st_32x32(tl + T_S, p);  // P over S: this row's scores are all in registers
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    // ---- O / L -> bf16 row of the output ----
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    uint32_t o[32], l;
    ptx::tmem_ld_32x32(tl + T_O, o);
    ptx::tmem_ld_32x1(tl + T_L, l);
    ptx::tmem_ld_wait();
    const float inv = 1.0f / __uint_as_float(l);
    __nv_bfloat16* dst = out + static_cast<size_t>(row_base + q0 + warp * 32 + lane) * ld_o + h * AHD;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      uint4 v;
      v.x = pack_bf16(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
      v.y = pack_bf16(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
      v.z = pack_bf16(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
      v.w = pack_bf16(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
      *reinterpret_cast<uint4*>(dst + 8 * c) = v;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, T_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Qwen2 prefill attention (causal, grouped-query, head_dim 64) on the same machinery  [transformers
// modeling_qwen2.py:149-184, 187-247].  T' is short (256 image tokens + prompt), so the tile is built the other way
// round: a CTA takes 16 consecutive queries of ALL query heads of one kv head (G = heads_q / heads_kv <= 8 heads x 16
// queries = up to 128 rows of one M = 128 MMA tile: row = head * 16 + query), i.e. K and V of the group are fetched
// once and no row of the tile lies beyond the causal diagonal's 64-key tile.  Key tiles of 64: tiles wholly below the
// diagonal need no mask, the last one masks key > query per element.  S / P / O in tensor memory as above
// (S [0,64) with P over it, O [64,128): 128 columns, 4 CTAs per SM); the row sums are added up by the softmax threads
// (no columns left for a ones-MMA accumulator).  RoPE has already been applied (qkv GEMM epilogue / rope kernel).
constexpr int CQ = 16;            // queries per CTA (per head of the group)
constexpr int CHD = 64;
constexpr int C_STAGES = 2;
constexpr int CQ_BYTES = 128 * CHD * 2;     // 16 KB: up to 8 heads x 16 queries
constexpr int CKV_BYTES = AKV * CHD * 2;    // 8 KB each for K and V
constexpr int C_SMEM = CQ_BYTES + C_STAGES * 2 * CKV_BYTES + 256 + 1024;
constexpr uint32_t TC_S = 0, TC_O = 64, TC_COLS = 128;

__global__ void __launch_bounds__(A_THREADS, A_MIN_CTAS)
attn_tc_causal_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                      const __grid_constant__ CUtensorMap tmap_v, __nv_bfloat16* __restrict__ out, int ld_o, int N,
                      int G, float scale_log2) {
  extern __shared__ uint8_t smem_attn[];
  const uint32_t base = (ptx::smem_u32(smem_attn) + 1023u) & ~1023u;
  const uint32_t s_q = base;
  const uint32_t s_kv = s_q + CQ_BYTES;
  const uint32_t bars = s_kv + C_STAGES * 2 * CKV_BYTES;
  const uint32_t q_full = bars;
  auto kv_full = [&](int s) { return bars + 8u + 8u * s; };
  auto kv_empty = [&](int s) { return bars + 8u + 8u * (C_STAGES + s); };
  const uint32_t s_full = bars + 8u + 8u * (2 * C_STAGES);
  const uint32_t p_full = s_full + 8u, o_full = s_full + 16u;
  const uint32_t tmem_ptr_smem = s_full + 24u;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x, g = blockIdx.y, b = blockIdx.z;
  const int q0 = qt * CQ;
  const int row_base = b * N;
  const int nkv = (q0 + CQ + AKV - 1) / AKV;  // key tiles up to and including the diagonal one

  if (warp == 5 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int s = 0; s < C_STAGES; ++s) {
      ptx::mbar_init(kv_full(s), 1);
      ptx::mbar_init(kv_empty(s), 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 4);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  // rows of the tile that no head of the group fills (G < 8): zero, so the MMAs never see stale bits as NaN / Inf
  for (int i = threadIdx.x; i < (128 - G * CQ) * (CHD * 2 / 16); i += A_THREADS)
    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(s_q + static_cast<uint32_t>(G * CQ) * 128u + 16u * i),
                 "r"(0u)
                 : "memory");
  if (warp == 4) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmap_q);
      ptx::prefetch_tmap(&tmap_k);
      ptx::prefetch_tmap(&tmap_v);
    }
    ptx::tmem_alloc(tmem_ptr_smem, TC_COLS);
    ptx::tmem_relinquish();
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_ptr_smem));
  pdl_sync();

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, static_cast<uint32_t>(G) * CQ * CHD * 2);
      for (int hh = 0; hh < G; ++hh)  // 16 queries of head g*G+hh -> tile rows [16 hh, 16 hh + 16)
        ptx::tma_load_2d(s_q + hh * CQ * CHD * 2, &tmap_q, (g * G + hh) * CHD, row_base + q0, q_full);
      int stage = 0;
      uint32_t phase = 0;
      for (int j = 0; j < nkv; ++j) {
        ptx::mbar_wait(kv_empty(stage), phase ^ 1u);
        ptx::mbar_arrive_expect_tx(kv_full(stage), 2 * CKV_BYTES);
        ptx::tma_load_2d(s_kv + stage * 2 * CKV_BYTES, &tmap_k, g * CHD, row_base + j * AKV, kv_full(stage));
        ptx::tma_load_2d(s_kv + stage * 2 * CKV_BYTES + CKV_BYTES, &tmap_v, g * CHD, row_base + j * AKV, kv_full(stage));
        if (++stage == C_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, AKV);
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16_bmn(128, CHD);
    const uint64_t dq = ptx::make_kmajor_sw128_desc(s_q);
    int stage = 0;
    uint32_t phase = 0;
    auto issue_qk = [&](int st) {
      const uint64_t dk = ptx::make_kmajor_sw128_desc(s_kv + st * 2 * CKV_BYTES);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < CHD / 16; ++k)
          ptx::umma_bf16(tmem_base + TC_S, dq + static_cast<uint64_t>(2 * k), dk + static_cast<uint64_t>(2 * k),
                         idesc_qk, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(kv_full(0), 0);
    ptx::tc_fence_after();
    issue_qk(0);
    for (int j = 0; j < nkv; ++j) {
      ptx::mbar_wait(p_full, j & 1);
      ptx::tc_fence_after();
      // V tile [key][64 dims] = MN-major B operand: 16 key rows per K step are 2048 B apart
      const uint64_t dv = ptx::make_kmajor_sw128_desc(s_kv + stage * 2 * CKV_BYTES + CKV_BYTES);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < AKV / 16; ++k)
          ptx::umma_bf16_ts(tmem_base + TC_O, tmem_base + TC_S + 8u * k, dv + static_cast<uint64_t>(128 * k), idesc_pv,
                            (j > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit(kv_empty(stage));
      }
      __syncwarp();
      if (++stage == C_STAGES) { stage = 0; phase ^= 1u; }
      if (j + 1 < nkv) {
        ptx::mbar_wait(kv_full(stage), phase);
        ptx::tc_fence_after();
        issue_qk(stage);
      } else {
        if (ptx::elect_one()) ptx::umma_commit(o_full);
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax: thread = tile row = (head hh, query qi) =====================
    const uint32_t tl = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    const int r = warp * 32 + lane;
    const int qi = q0 + (r & (CQ - 1));       // this row's query position
    const f32x2 c2 = splat2(scale_log2);
    float mx = -INFINITY, lsum = 0.f;
    for (int j = 0; j < nkv; ++j) {
      ptx::mbar_wait(s_full, j & 1);
      ptx::tc_fence_after();
      uint32_t s0[32], s1[32];
      ptx::tmem_ld_32x32(tl + TC_S, s0);
      ptx::tmem_ld_32x32(tl + TC_S + 32u, s1);
      ptx::tmem_ld_wait();
      if (j == nkv - 1) {  // the diagonal tile: keys beyond this row's query are masked
        const int lim = qi - j * AKV;  // last admissible column
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i > lim) s0[i] = 0xff800000u;       // -inf
          if (i + 32 > lim) s1[i] = 0xff800000u;
        }
      }
      float t0 = __uint_as_float(s0[0]), t1 = __uint_as_float(s0[1]), t2 = __uint_as_float(s1[0]),
            t3 = __uint_as_float(s1[1]);
#pragma unroll
      for (int i = 2; i < 32; i += 2) {
        t0 = fmaxf(t0, __uint_as_float(s0[i]));
        t1 = fmaxf(t1, __uint_as_float(s0[i + 1]));
        t2 = fmaxf(t2, __uint_as_float(s1[i]));
        t3 = fmaxf(t3, __uint_as_float(s1[i + 1]));
      }
      const float tm = fmaxf(fmaxf(t0, t1), fmaxf(t2, t3)) * scale_log2;
      const bool need = tm > mx + LAZY_LOG2;
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        const float f = need ? ex2_approx(mx - tm) : 1.0f;
        lsum *= f;
#pragma unroll 1
        for (int c = 0; c < CHD; c += 8) {
          uint32_t o[8];
          ptx::tmem_ld_32x8(tl + TC_O + c, o);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
          ptx::tmem_st_32x8(tl + TC_O + c, o);
        }
      }
      if (need) mx = tm;
      const f32x2 nm2 = splat2(-mx);
      uint32_t p[32];
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a0, a1, b0, b1;
        upk2(fma2(pk2(__uint_as_float(s0[2 * i]), __uint_as_float(s0[2 * i + 1])), c2, nm2), a0, a1);
        upk2(fma2(pk2(__uint_as_float(s1[2 * i]), __uint_as_float(s1[2 * i + 1])), c2, nm2), b0, b1);
        a0 = ex2_approx(a0); a1 = ex2_approx(a1); b0 = ex2_approx(b0); b1 = ex2_approx(b1);
        l0 += a0 + a1;
        l1 += b0 + b1;
        p[i] = pack_bf16(a0, a1);
        p[16 + i] = pack_bf16(b0, b1);
      }
      lsum += l0 + l1;
      ptx::tmem_st_32x32(tl + TC_S, p);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const int hh = r / CQ;
    const bool live = hh < G && qi < N;
    const float inv = 1.0f / lsum;
    __nv_bfloat16* dst = out + static_cast<size_t>(row_base + qi) * ld_o + (g * G + hh) * CHD;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t o[32];
      ptx::tmem_ld_32x32(tl + TC_O + 32u * half, o);
      ptx::tmem_ld_wait();
      if (live) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint4 v;
          v.x = pack_bf16(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv);
          v.y = pack_bf16(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv);
          v.z = pack_bf16(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv);
          v.w = pack_bf16(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + 32 * half + 8 * c) = v;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TC_COLS);
  }
}

// column slice [rows, cols] (row pitch ld elements) of the fused qkv buffer: box = 64 rows x 32 columns, 64-byte swizzle
int make_tmap_slice(CUtensorMap* out, const void* ptr, long long rows, long long cols, long long ld,
                    unsigned box_cols = 32u, unsigned box_rows = 64u) {
  TmaEncodeTiledFn fn = tma_encode_fn();
  FVLA_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  FVLA_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15u) == 0 && (ld * 2) % 16 == 0,
               "attention (tcgen05): q/k/v slices must be 16-byte aligned with a 16-byte multiple pitch");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1u, 1u};
  // rows of 32 bf16 land under the 64-byte swizzle, rows of 64 bf16 under the 128-byte one (the UMMA descriptors' modes)
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32u ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (attention) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 1;
  }
  return 0;
}

}  // namespace

namespace {
bool tc_common_ok(const AttnArgs& a) {
  return a.rope_cos == nullptr && a.heads_kv > 0 && a.heads_q % a.heads_kv == 0 && a.ld_qkv % 8 == 0 &&
         a.ld_o % 8 == 0 && (reinterpret_cast<uintptr_t>(a.q) & 15u) == 0 &&
         (reinterpret_cast<uintptr_t>(a.k) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.v) & 15u) == 0 &&
         (reinterpret_cast<uintptr_t>(a.o) & 15u) == 0;
}
bool tc_causal_shape(const AttnArgs& a) {  // Qwen2 prefill: causal, head_dim 64, a kv group of at most 8 query heads
  return a.causal && a.head_dim == CHD && a.heads_q / a.heads_kv <= 8 && a.N >= 1;
}
}  // namespace

bool attention_tc_supported(const AttnArgs& a) {
  static const bool on = std::getenv("FVLA_DISABLE_TC_ATTN") == nullptr;  // A/B switch for profiling
  if (!on || !tc_common_ok(a)) return false;
  if (tc_causal_shape(a)) return true;
  return a.head_dim == AHD && !a.causal && a.N >= AQ && a.N % AQ == 0;
}

int attention_tc(const AttnArgs& a, cudaStream_t stream) {
  FVLA_REQUIRE(attention_tc_supported(a), "attention_tc: unsupported shape");
  if (tc_causal_shape(a)) {
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_tc_causal_kernel), C_SMEM)) return rc;
    const long long rows = static_cast<long long>(a.B) * a.N;
    const int G = a.heads_q / a.heads_kv;
    CUtensorMap tq, tk, tv;
    if (int rc = make_tmap_slice(&tq, a.q, rows, static_cast<long long>(a.heads_q) * CHD, a.ld_qkv, 64u, CQ)) return rc;
    if (int rc = make_tmap_slice(&tk, a.k, rows, static_cast<long long>(a.heads_kv) * CHD, a.ld_qkv, 64u, AKV)) return rc;
    if (int rc = make_tmap_slice(&tv, a.v, rows, static_cast<long long>(a.heads_kv) * CHD, a.ld_qkv, 64u, AKV)) return rc;
    dim3 grid(ceil_div(a.N, CQ), a.heads_kv, a.B);
    FVLA_CUDA_CHECK(launch_pdl(attn_tc_causal_kernel, grid, dim3(A_THREADS), C_SMEM, stream, tq, tk, tv,
                               static_cast<__nv_bfloat16*>(a.o), a.ld_o, a.N, G, a.scale * 1.4426950408889634f));
    return 0;
  }
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(attn_tc_kernel), A_SMEM)) return rc;
  const long long rows = static_cast<long long>(a.B) * a.N;
  CUtensorMap tq, tk, tv;
  if (int rc = make_tmap_slice(&tq, a.q, rows, static_cast<long long>(a.heads_q) * AHD, a.ld_qkv)) return rc;
  if (int rc = make_tmap_slice(&tk, a.k, rows, static_cast<long long>(a.heads_kv) * AHD, a.ld_qkv)) return rc;
  if (int rc = make_tmap_slice(&tv, a.v, rows, static_cast<long long>(a.heads_kv) * AHD, a.ld_qkv)) return rc;
  dim3 grid(a.N / AQ, a.heads_q, a.B);
  FVLA_CUDA_CHECK(launch_pdl(attn_tc_kernel, grid, dim3(A_THREADS), A_SMEM, stream, tq, tk, tv,
                             static_cast<__nv_bfloat16*>(a.o), a.ld_o, a.N, a.heads_q / a.heads_kv,
                             a.scale * 1.4426950408889634f));
  return 0;
}

}  // namespace fvla
