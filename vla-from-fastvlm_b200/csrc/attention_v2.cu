// Flash attention, second version (bf16, no rotary: Qwen2's RoPE is applied in place by rope_inplace).
//
// Differences from flash_attn_bf16_kernel (attention.cu), all aimed at instructions per score —
// with head_dim 32 a score gets only 128 tensor-core flops, so the softmax bookkeeping dominates:
//   * 128 queries per CTA (8 warps x 16 rows): K/V tiles are staged half as often per query
//   * K/V tiles double-buffered with cp.async (zero-fill past the sequence end)
//   * Q and K fragments via ldmatrix.x4 instead of scalar 32-bit shared loads
//   * the softmax scale is folded into one FFMA feeding ex2: p = 2^(s*c - m*c)
//   * masking code only runs on tiles that need it (sequence tail, causal diagonal); warps whose rows
//     lie entirely before a causal tile skip it
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

constexpr int BKV2 = 64;

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t pk2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void cpa16(uint32_t dst, const void* src, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(bytes) : "memory");
}

// rows x HD tile -> padded smem (row pitch HD+8 elements), 16-byte cp.async, zero fill beyond n_tok
template <int HD, int ROWS, int THREADS>
__device__ __forceinline__ void stage_async(uint32_t dst, const __nv_bfloat16* src, int ld, int tok0, int n_tok) {
  constexpr int VPR = HD / 8;
  constexpr int LDSB = (HD + 8) * 2;
  for (int i = threadIdx.x; i < ROWS * VPR; i += THREADS) {
    const int r = i / VPR, c = i % VPR;
    const bool ok = tok0 + r < n_tok;
    const __nv_bfloat16* p = ok ? src + static_cast<size_t>(tok0 + r) * ld + c * 8 : src;
    cpa16(dst + r * LDSB + c * 16, p, ok ? 16 : 0);
  }
}

// NWARPS x 16 queries per CTA: 8 warps for long sequences (K/V tiles staged half as often per query), 4 warps for
// the short causal Qwen2 prefill (T' = 272: 128-query CTAs would leave 29 % of their rows empty)
template <int HD, bool CAUSAL, int NWARPS>
__global__ void __launch_bounds__(32 * NWARPS, (NWARPS == 8 ? (HD <= 32 ? 3 : (HD <= 64 ? 2 : 1)) : (HD <= 64 ? 4 : 2)))
flash_attn_v2_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                     const __nv_bfloat16* __restrict__ v, int ld, __nv_bfloat16* __restrict__ o, int ld_o,
                     int N, int heads_q, int heads_kv, float scale_log2e) {
  constexpr int BQ2 = 16 * NWARPS, NT = 32 * NWARPS;
  constexpr int LDS = HD + 8;          // elements
  constexpr int LDSB = LDS * 2;        // bytes
  constexpr int KV_TILE_B = BKV2 * LDSB;
  extern __shared__ __align__(16) uint8_t smem_a2[];
  const uint32_t sQ = static_cast<uint32_t>(__cvta_generic_to_shared(smem_a2));
  const uint32_t sK = sQ + BQ2 * LDSB;            // 2 stages
  const uint32_t sV = sK + 2 * KV_TILE_B;         // 2 stages

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (heads_q / heads_kv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const size_t row0 = static_cast<size_t>(b) * N;
  const __nv_bfloat16* qp = q + row0 * ld + h * HD;
  const __nv_bfloat16* kp = k + row0 * ld + hk * HD;
  const __nv_bfloat16* vp = v + row0 * ld + hk * HD;

  const int q0 = qb * BQ2;
  int kv_blocks = (N + BKV2 - 1) / BKV2;
  if (CAUSAL) kv_blocks = min(kv_blocks, (q0 + BQ2 + BKV2 - 1) / BKV2);

  stage_async<HD, BQ2, NT>(sQ, qp, ld, q0, N);
  stage_async<HD, BKV2, NT>(sK, kp, ld, 0, N);
  stage_async<HD, BKV2, NT>(sV, vp, ld, 0, N);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // Q fragments (A operand) of this warp's 16 rows
  uint32_t qf[HD / 16][4];
  {
    const uint32_t base = sQ + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDSB + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) ldsm_x4(qf[ks], base + ks * 32);
  }

  float oacc[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) { oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};  // running max of the RAW scores
  float l_run[2] = {0.f, 0.f};
  const int qrow = q0 + warp * 16 + g;      // this thread's rows: qrow and qrow + 8
  const int wrow_max = q0 + warp * 16 + 15;

  for (int kb = 0; kb < kv_blocks; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < kv_blocks) {  // prefetch the next tile into the other stage
      stage_async<HD, BKV2, NT>(sK + (st ^ 1) * KV_TILE_B, kp, ld, (kb + 1) * BKV2, N);
      stage_async<HD, BKV2, NT>(sV + (st ^ 1) * KV_TILE_B, vp, ld, (kb + 1) * BKV2, N);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");

    const int key0 = kb * BKV2;
    const bool skip = CAUSAL && key0 > wrow_max;  // every key of the tile lies after every row of the warp
    if (!skip) {
      const uint32_t kt = sK + st * KV_TILE_B, vt = sV + st * KV_TILE_B;
      // ---- S = Q K^T ----
      float s[BKV2 / 8][4];
#pragma unroll
      for (int j = 0; j < BKV2 / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
      for (int j = 0; j < BKV2 / 8; j += 2) {
        // matrices: (keys j*8.., dims ks*16..+7), (same keys, +8), (keys (j+1)*8.., ..), (.., +8)
        const uint32_t kaddr = kt + ((j + (lane >> 4)) * 8 + (lane & 7)) * LDSB + ((lane >> 3) & 1) * 16;
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
          uint32_t kf[4];
          ldsm_x4(kf, kaddr + ks * 32);
          mma16816(s[j], qf[ks], kf[0], kf[1]);
          mma16816(s[j + 1], qf[ks], kf[2], kf[3]);
        }
      }
      // ---- mask (only where needed) ----
      const bool need_mask = (key0 + BKV2 > N) || (CAUSAL && key0 + BKV2 - 1 > q0 + warp * 16);
      if (need_mask) {
#pragma unroll
        for (int j = 0; j < BKV2 / 8; ++j) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int key = key0 + j * 8 + 2 * tig + (e & 1);
            const int qi = qrow + (e >> 1) * 8;
            const bool dead = key >= N || (CAUSAL && key > qi);
            if (dead) s[j][e] = -INFINITY;
          }
        }
      }
      // ---- online softmax ----
      float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
      for (int j = 0; j < BKV2 / 8; ++j) {
        mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
        mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
      }
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      }
      float ms[2], corr[2];
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const float msafe = mx[r] == -INFINITY ? 0.f : mx[r];
        ms[r] = msafe * scale_log2e;
        corr[r] = ex2f(m_run[r] * scale_log2e - ms[r]);  // m_run = -inf -> 0
        m_run[r] = mx[r];
      }
      float lsum[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < BKV2 / 8; ++j) {
        s[j][0] = ex2f(fmaf(s[j][0], scale_log2e, -ms[0]));
        s[j][1] = ex2f(fmaf(s[j][1], scale_log2e, -ms[0]));
        s[j][2] = ex2f(fmaf(s[j][2], scale_log2e, -ms[1]));
        s[j][3] = ex2f(fmaf(s[j][3], scale_log2e, -ms[1]));
        lsum[0] += s[j][0] + s[j][1];
        lsum[1] += s[j][2] + s[j][3];
      }
      l_run[0] = fmaf(l_run[0], corr[0], lsum[0]);
      l_run[1] = fmaf(l_run[1], corr[1], lsum[1]);
#pragma unroll
      for (int j = 0; j < HD / 8; ++j) {
        oacc[j][0] *= corr[0]; oacc[j][1] *= corr[0];
        oacc[j][2] *= corr[1]; oacc[j][3] *= corr[1];
      }
      // ---- O += P V ----
#pragma unroll
      for (int kk = 0; kk < BKV2 / 16; ++kk) {
        uint32_t pa[4];
        pa[0] = pk2(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pk2(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pk2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pk2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const uint32_t vaddr = vt + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDSB + (lane >> 4) * 16;
#pragma unroll
        for (int jn = 0; jn < HD / 8; jn += 2) {
          uint32_t vb[4];
          ldsm_x4_t(vb, vaddr + jn * 16);
          mma16816(oacc[jn], pa, vb[0], vb[1]);
          mma16816(oacc[jn + 1], pa, vb[2], vb[3]);
        }
      }
    }
    // next tile landed; everyone is done reading this stage before it is overwritten next iteration
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }

  // ---- normalise and store ----
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  __nv_bfloat16* op = o + row0 * ld_o + h * HD;
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    const int col = j * 8 + 2 * tig;
    if (qrow < N)
      *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(qrow) * ld_o + col) =
          pk2(oacc[j][0] * inv0, oacc[j][1] * inv0);
    if (qrow + 8 < N)
      *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(qrow + 8) * ld_o + col) =
          pk2(oacc[j][2] * inv1, oacc[j][3] * inv1);
  }
}

// Grouped-query variant for the short causal Qwen2 prefill: a CTA owns 16 queries of ONE kv head and runs one warp
// per query head of that group (7 for Qwen2-0.5B/7B, 6 for 1.5B), so a K/V tile is staged once for the whole group
// instead of once per query head (registers cap an SM at ~16 warps either way).  Same fragment code as above with
// every warp on the same 16 rows and its own head.
template <int HD>
__global__ void __launch_bounds__(256, (HD <= 64 ? 2 : 1))
flash_attn_gqa_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                      const __nv_bfloat16* __restrict__ v, int ld, __nv_bfloat16* __restrict__ o, int ld_o,
                      int N, int group, float scale_log2e) {
  constexpr int LDS = HD + 8;
  constexpr int LDSB = LDS * 2;
  constexpr int KV_TILE_B = BKV2 * LDSB;
  constexpr int VPR = HD / 8;
  extern __shared__ __align__(16) uint8_t smem_a3[];
  const uint32_t sK = static_cast<uint32_t>(__cvta_generic_to_shared(smem_a3));  // 2 stages
  const uint32_t sV = sK + 2 * KV_TILE_B;                                         // 2 stages
  const uint32_t sQ = sV + 2 * KV_TILE_B;                                         // [group][16 rows]

  const int qb = blockIdx.x, hk = blockIdx.y, b = blockIdx.z;
  const int nthr = blockDim.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const int h = hk * group + warp;  // this warp's query head
  const size_t row0 = static_cast<size_t>(b) * N;
  const __nv_bfloat16* kp = k + row0 * ld + hk * HD;
  const __nv_bfloat16* vp = v + row0 * ld + hk * HD;
  const int q0 = qb * 16;
  const int kv_blocks = min((N + BKV2 - 1) / BKV2, (q0 + 16 + BKV2 - 1) / BKV2);

  auto stage_kv = [&](int st, int tok0) {
    for (int i = threadIdx.x; i < BKV2 * VPR; i += nthr) {
      const int r = i / VPR, c = i % VPR;
      const bool ok = tok0 + r < N;
      const size_t off = ok ? static_cast<size_t>(tok0 + r) * ld + c * 8 : 0;
      cpa16(sK + st * KV_TILE_B + r * LDSB + c * 16, kp + off, ok ? 16 : 0);
      cpa16(sV + st * KV_TILE_B + r * LDSB + c * 16, vp + off, ok ? 16 : 0);
    }
  };
  {  // Q: 16 rows of each head of the group (adjacent heads are adjacent columns of the qkv row)
    const __nv_bfloat16* qg = q + row0 * ld + hk * group * HD;
    for (int i = threadIdx.x; i < group * 16 * VPR; i += nthr) {
      const int c = i % VPR, r = (i / VPR) % 16, w = i / (VPR * 16);
      const bool ok = q0 + r < N;
      const __nv_bfloat16* p = ok ? qg + static_cast<size_t>(q0 + r) * ld + w * HD + c * 8 : qg;
      cpa16(sQ + (w * 16 + r) * LDSB + c * 16, p, ok ? 16 : 0);
    }
  }
  stage_kv(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  uint32_t qf[HD / 16][4];
  {
    const uint32_t base = sQ + (warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDSB + (lane >> 4) * 16;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) ldsm_x4(qf[ks], base + ks * 32);
  }
  float oacc[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) { oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};
  const int qrow = q0 + g;  // this thread's rows: qrow and qrow + 8

  for (int kb = 0; kb < kv_blocks; ++kb) {
    const int st = kb & 1;
    if (kb + 1 < kv_blocks) stage_kv(st ^ 1, (kb + 1) * BKV2);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const int key0 = kb * BKV2;
    const uint32_t kt = sK + st * KV_TILE_B, vt = sV + st * KV_TILE_B;
    float s[BKV2 / 8][4];
#pragma unroll
    for (int j = 0; j < BKV2 / 8; ++j) { s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f; }
#pragma unroll
    for (int j = 0; j < BKV2 / 8; j += 2) {
      const uint32_t kaddr = kt + ((j + (lane >> 4)) * 8 + (lane & 7)) * LDSB + ((lane >> 3) & 1) * 16;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t kf[4];
        ldsm_x4(kf, kaddr + ks * 32);
        mma16816(s[j], qf[ks], kf[0], kf[1]);
        mma16816(s[j + 1], qf[ks], kf[2], kf[3]);
      }
    }
    if (key0 + BKV2 > N || key0 + BKV2 - 1 > q0) {  // sequence tail or causal diagonal
#pragma unroll
      for (int j = 0; j < BKV2 / 8; ++j) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = key0 + j * 8 + 2 * tig + (e & 1);
          const int qi = qrow + (e >> 1) * 8;
          if (key >= N || key > qi) s[j][e] = -INFINITY;
        }
      }
    }
    float mx[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int j = 0; j < BKV2 / 8; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(s[j][0], s[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[j][2], s[j][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float ms[2], corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float msafe = mx[r] == -INFINITY ? 0.f : mx[r];
      ms[r] = msafe * scale_log2e;
      corr[r] = ex2f(m_run[r] * scale_log2e - ms[r]);
      m_run[r] = mx[r];
    }
    float lsum[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < BKV2 / 8; ++j) {
      s[j][0] = ex2f(fmaf(s[j][0], scale_log2e, -ms[0]));
      s[j][1] = ex2f(fmaf(s[j][1], scale_log2e, -ms[0]));
      s[j][2] = ex2f(fmaf(s[j][2], scale_log2e, -ms[1]));
      s[j][3] = ex2f(fmaf(s[j][3], scale_log2e, -ms[1]));
      lsum[0] += s[j][0] + s[j][1];
      lsum[1] += s[j][2] + s[j][3];
    }
    l_run[0] = fmaf(l_run[0], corr[0], lsum[0]);
    l_run[1] = fmaf(l_run[1], corr[1], lsum[1]);
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
      oacc[j][0] *= corr[0]; oacc[j][1] *= corr[0];
      oacc[j][2] *= corr[1]; oacc[j][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < BKV2 / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pk2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pk2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pk2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pk2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
      const uint32_t vaddr = vt + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDSB + (lane >> 4) * 16;
#pragma unroll
      for (int jn = 0; jn < HD / 8; jn += 2) {
        uint32_t vb[4];
        ldsm_x4_t(vb, vaddr + jn * 16);
        mma16816(oacc[jn], pa, vb[0], vb[1]);
        mma16816(oacc[jn + 1], pa, vb[2], vb[3]);
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = l_run[0] > 0.f ? 1.f / l_run[0] : 0.f;
  const float inv1 = l_run[1] > 0.f ? 1.f / l_run[1] : 0.f;
  __nv_bfloat16* op = o + row0 * ld_o + h * HD;
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) {
    const int col = j * 8 + 2 * tig;
    if (qrow < N)
      *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(qrow) * ld_o + col) = pk2(oacc[j][0] * inv0, oacc[j][1] * inv0);
    if (qrow + 8 < N)
      *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(qrow + 8) * ld_o + col) = pk2(oacc[j][2] * inv1, oacc[j][3] * inv1);
  }
}

template <int HD>
int launch_gqa(const AttnArgs& a, cudaStream_t stream) {
  auto kfn = flash_attn_gqa_kernel<HD>;
  const int group = a.heads_q / a.heads_kv;
  const int smem = (4 * BKV2 + 8 * 16) * (HD + 8) * 2;  // Q sized for the largest group
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), smem)) return rc;
  dim3 grid(ceil_div(a.N, 16), a.heads_kv, a.B);
  kfn<<<grid, 32 * group, smem, stream>>>(static_cast<const __nv_bfloat16*>(a.q), static_cast<const __nv_bfloat16*>(a.k),
                                          static_cast<const __nv_bfloat16*>(a.v), a.ld_qkv, static_cast<__nv_bfloat16*>(a.o),
                                          a.ld_o, a.N, group, a.scale * 1.4426950408889634f);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <int HD, bool CAUSAL, int NWARPS>
int launch_v2(const AttnArgs& a, cudaStream_t stream) {
  auto kfn = flash_attn_v2_kernel<HD, CAUSAL, NWARPS>;
  constexpr int BQ2 = 16 * NWARPS;
  constexpr int SMEM = (BQ2 + 4 * BKV2) * (HD + 8) * 2;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), SMEM)) return rc;
  dim3 grid(ceil_div(a.N, BQ2), a.heads_q, a.B);
  kfn<<<grid, 32 * NWARPS, SMEM, stream>>>(static_cast<const __nv_bfloat16*>(a.q), static_cast<const __nv_bfloat16*>(a.k),
                                   static_cast<const __nv_bfloat16*>(a.v), a.ld_qkv,
                                   static_cast<__nv_bfloat16*>(a.o), a.ld_o, a.N, a.heads_q, a.heads_kv,
                                   a.scale * 1.4426950408889634f);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace

bool attention_v2_supported(const AttnArgs& a) {
  return a.rope_cos == nullptr && (a.head_dim == 32 || a.head_dim == 64 || a.head_dim == 128);
}

int attention_v2(const AttnArgs& a, cudaStream_t stream) {
  const bool small = a.causal && a.N < 512;  // the short causal prefill
  // grouped-query CTAs (one warp per query head of a kv group) when the group is 2..8 heads
  static const bool gqa_on = std::getenv("FVLA_DISABLE_GQA_ATTN") == nullptr;  // A/B switch
  if (gqa_on && small && a.heads_kv > 0 && a.heads_q % a.heads_kv == 0 && a.heads_q / a.heads_kv >= 2 &&
      a.heads_q / a.heads_kv <= 8 && (a.head_dim == 64 || a.head_dim == 128))
    return a.head_dim == 64 ? launch_gqa<64>(a, stream) : launch_gqa<128>(a, stream);
  if (a.head_dim == 32) {
    if (small) return a.causal ? launch_v2<32, true, 4>(a, stream) : launch_v2<32, false, 4>(a, stream);
    return a.causal ? launch_v2<32, true, 8>(a, stream) : launch_v2<32, false, 8>(a, stream);
  }
  if (a.head_dim == 64) {
    if (small) return a.causal ? launch_v2<64, true, 4>(a, stream) : launch_v2<64, false, 4>(a, stream);
    return a.causal ? launch_v2<64, true, 8>(a, stream) : launch_v2<64, false, 8>(a, stream);
  }
  if (small) return a.causal ? launch_v2<128, true, 4>(a, stream) : launch_v2<128, false, 4>(a, stream);
  return a.causal ? launch_v2<128, true, 8>(a, stream) : launch_v2<128, false, 8>(a, stream);
}

}  // namespace fvla
