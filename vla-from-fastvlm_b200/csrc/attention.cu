// Fused attention kernels.
//  * flash_attn_bf16_kernel: FastViTHD MHSA (non-causal, head_dim 32, N = 1024 / 256) and Qwen2
//    GQA prefill (causal, head_dim 64 / 128, rotary embedding applied while staging Q and K).
//    One CTA = 64 queries of one (sample, head); 4 warps x 16 query rows; online softmax in fp32;
//    S = Q K^T and O += P V on bf16 tensor-core MMAs (m16n8k16); the N x N matrix never leaves
//    registers.
//  * attn_simt_kernel: straightforward fp32 kernel (one warp per query) used by the fp32 parity
//    mode and as the on-GPU cross-check of the flash kernel in the tests.
// q/k/v are column slices of the fused qkv GEMM output, so they share one row pitch.
#include "common.cuh"
#include "kernels.h"

namespace fvla {
namespace {

constexpr int BQ = 64;   // queries per CTA
constexpr int BKV = 64;  // keys per inner iteration

__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Stage a [rows x HD] tile (row r <-> token tok0 + r) into padded smem, optionally rotating
// (rotate-half RoPE, transformers/models/qwen2/modeling_qwen2.py: apply_rotary_pos_emb).
template <int HD, bool ROPE>
__device__ __forceinline__ void stage_tile(__nv_bfloat16* dst, const __nv_bfloat16* src, int ld,
                                           int tok0, int n_tok, const float* rope_cos,
                                           const float* rope_sin) {
  constexpr int LDS = HD + 8;
  if constexpr (!ROPE) {
    constexpr int VPR = HD / 8;  // 16-byte vectors per row
    for (int i = threadIdx.x; i < 64 * VPR; i += blockDim.x) {
      const int r = i / VPR, c = (i % VPR) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (tok0 + r < n_tok)
        v = __ldg(reinterpret_cast<const uint4*>(src + static_cast<size_t>(tok0 + r) * ld + c));
      *reinterpret_cast<uint4*>(dst + r * LDS + c) = v;
    }
  } else {
    constexpr int HALF = HD / 2;
    constexpr int VPR = HALF / 8;
    for (int i = threadIdx.x; i < 64 * VPR; i += blockDim.x) {
      const int r = i / VPR, c = (i % VPR) * 8;
      Vec8<__nv_bfloat16> lo, hi;
      const int tok = tok0 + r;
      if (tok < n_tok) {
        const __nv_bfloat16* p = src + static_cast<size_t>(tok) * ld + c;
        lo.load(p);
        hi.load(p + HALF);
        const float* cs = rope_cos + static_cast<size_t>(tok) * HALF + c;
        const float* sn = rope_sin + static_cast<size_t>(tok) * HALF + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float co = __ldg(cs + j), si = __ldg(sn + j);
          const float a = lo.v[j], b = hi.v[j];
          lo.v[j] = a * co - b * si;
          hi.v[j] = b * co + a * si;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) { lo.v[j] = 0.f; hi.v[j] = 0.f; }
      }
      lo.store(dst + r * LDS + c);
      hi.store(dst + r * LDS + c + HALF);
    }
  }
}

template <int HD, bool CAUSAL, bool ROPE>
__global__ void __launch_bounds__(128)
flash_attn_bf16_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                       const __nv_bfloat16* __restrict__ v, int ld, __nv_bfloat16* __restrict__ o,
                       int ld_o, int N, int heads_q, int heads_kv, float scale_log2e,
                       const float* __restrict__ rope_cos, const float* __restrict__ rope_sin) {
  constexpr int LDS = HD + 8;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* Qs = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* Ks = Qs + BQ * LDS;
  __nv_bfloat16* Vs = Ks + BKV * LDS;

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int hk = h / (heads_q / heads_kv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, tig = lane & 3;
  const size_t row0 = static_cast<size_t>(b) * N;
  const __nv_bfloat16* qp = q + row0 * ld + h * HD;
  const __nv_bfloat16* kp = k + row0 * ld + hk * HD;
  const __nv_bfloat16* vp = v + row0 * ld + hk * HD;

  stage_tile<HD, ROPE>(Qs, qp, ld, qb * BQ, N, rope_cos, rope_sin);
  __syncthreads();

  // Q fragments for this warp's 16 rows
  uint32_t qf[HD / 16][4];
  {
    const __nv_bfloat16* base = Qs + (warp * 16) * LDS;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      qf[ks][0] = *reinterpret_cast<const uint32_t*>(base + g * LDS + ks * 16 + 2 * tig);
      qf[ks][1] = *reinterpret_cast<const uint32_t*>(base + (g + 8) * LDS + ks * 16 + 2 * tig);
      qf[ks][2] = *reinterpret_cast<const uint32_t*>(base + g * LDS + ks * 16 + 8 + 2 * tig);
      qf[ks][3] = *reinterpret_cast<const uint32_t*>(base + (g + 8) * LDS + ks * 16 + 8 + 2 * tig);
    }
  }

  float oacc[HD / 8][4];
#pragma unroll
  for (int j = 0; j < HD / 8; ++j) { oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f; }
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  const int q_row_lo = qb * BQ + warp * 16 + g;  // global query index of this thread's first row
  const int kv_blocks = CAUSAL ? (qb + 1) : (N + BKV - 1) / BKV;

  for (int kb = 0; kb < kv_blocks; ++kb) {
    __syncthreads();  // previous iteration's reads of Ks/Vs are done
    stage_tile<HD, ROPE>(Ks, kp, ld, kb * BKV, N, rope_cos, rope_sin);
    stage_tile<HD, false>(Vs, vp, ld, kb * BKV, N, nullptr, nullptr);
    __syncthreads();

    // ---- S = Q K^T (16 x 64 per warp) ----
    float s[BKV / 8][4];
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      const __nv_bfloat16* krow = Ks + (j * 8 + g) * LDS + 2 * tig;
#pragma unroll
      for (int ks = 0; ks < HD / 16; ++ks) {
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(krow + ks * 16);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(krow + ks * 16 + 8);
        mma_bf16_16816(s[j], qf[ks], b0, b1);
      }
    }
    // ---- scale, mask, online softmax ----
    float m_new[2] = {m_run[0], m_run[1]};
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb * BKV + j * 8 + 2 * tig + (e & 1);
        const int qi = q_row_lo + (e >> 1) * 8;
        float val = s[j][e] * scale_log2e;
        bool masked = key >= N;
        if (CAUSAL) masked = masked || key > qi;
        val = masked ? -INFINITY : val;
        s[j][e] = val;
        m_new[e >> 1] = fmaxf(m_new[e >> 1], val);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 1));
      m_new[r] = fmaxf(m_new[r], __shfl_xor_sync(0xffffffffu, m_new[r], 2));
    }
    float corr[2], msafe[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      msafe[r] = m_new[r] == -INFINITY ? 0.f : m_new[r];
      corr[r] = exp2f(m_run[r] - msafe[r]);  // m_run = -inf -> 0
      m_run[r] = m_new[r];
      l_run[r] *= corr[r];
    }
    float lsum[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < BKV / 8; ++j) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float pv = exp2f(s[j][e] - msafe[e >> 1]);
        s[j][e] = pv;
        lsum[e >> 1] += pv;
      }
    }
    l_run[0] += lsum[0];
    l_run[1] += lsum[1];
#pragma unroll
    for (int j = 0; j < HD / 8; ++j) {
      oacc[j][0] *= corr[0]; oacc[j][1] *= corr[0];
      oacc[j][2] *= corr[1]; oacc[j][3] *= corr[1];
    }
    // ---- O += P V ----
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t pa[4];
      pa[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int jn = 0; jn < HD / 8; jn += 2) {
        // four 8x8 blocks of V: (keys kk*16 + {0..7, 8..15}) x (dims jn*8 + {0..7, 8..15})
        const int vrow = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
        const int vcol = jn * 8 + (lane >> 4) * 8;
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, static_cast<uint32_t>(__cvta_generic_to_shared(Vs + vrow * LDS + vcol)));
        mma_bf16_16816(oacc[jn], pa, vb[0], vb[1]);
        mma_bf16_16816(oacc[jn + 1], pa, vb[2], vb[3]);
      }
    }
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
    if (q_row_lo < N)
      *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(q_row_lo) * ld_o + col) =
          pack2(oacc[j][0] * inv0, oacc[j][1] * inv0);
    if (q_row_lo + 8 < N)
      *reinterpret_cast<uint32_t*>(op + static_cast<size_t>(q_row_lo + 8) * ld_o + col) =
          pack2(oacc[j][2] * inv1, oacc[j][3] * inv1);
  }
}

// ---------------------------------------------------------------------------------------------
// SIMT attention: one warp per query. Scores for all keys live in shared memory.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float rope_elem(const T* x, int i, int hd, const float* cs,
                                           const float* sn) {
  // x: one head row; cs/sn: this token's [hd/2] tables (null = no rotation)
  const float xi = to_f32(x[i]);
  if (cs == nullptr) return xi;
  const int half = hd >> 1;
  if (i < half) return xi * cs[i] - to_f32(x[i + half]) * sn[i];
  return xi * cs[i - half] + to_f32(x[i - half]) * sn[i - half];
}

template <typename T>
__global__ void __launch_bounds__(128)
attn_simt_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v, int ld,
                 T* __restrict__ o, int ld_o, int N, int heads_q, int heads_kv, int hd, float scale,
                 int causal, const float* __restrict__ rope_cos, const float* __restrict__ rope_sin) {
  extern __shared__ float sm[];  // per warp: [hd] query + [N] scores
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qi = blockIdx.x * 4 + warp;
  const int h = blockIdx.y, b = blockIdx.z;
  if (qi >= N) return;
  const int hk = h / (heads_q / heads_kv);
  float* qs = sm + static_cast<size_t>(warp) * (hd + N);
  float* sc = qs + hd;
  const size_t row0 = static_cast<size_t>(b) * N;
  const int half = hd >> 1;
  const T* qrow = q + (row0 + qi) * ld + h * hd;
  for (int i = lane; i < hd; i += 32)
    qs[i] = rope_elem(qrow, i, hd, rope_cos ? rope_cos + static_cast<size_t>(qi) * half : nullptr,
                      rope_sin ? rope_sin + static_cast<size_t>(qi) * half : nullptr) * scale;
  __syncwarp();
  const int n_keys = causal ? (qi + 1) : N;
  float mx = -INFINITY;
  for (int j = lane; j < n_keys; j += 32) {
    const T* krow = k + (row0 + j) * ld + hk * hd;
    const float* cs = rope_cos ? rope_cos + static_cast<size_t>(j) * half : nullptr;
    const float* sn = rope_sin ? rope_sin + static_cast<size_t>(j) * half : nullptr;
    float a = 0.f;
    for (int i = 0; i < hd; ++i) a = fmaf(qs[i], rope_elem(krow, i, hd, cs, sn), a);
    sc[j] = a;
    mx = fmaxf(mx, a);
  }
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < n_keys; j += 32) {
    const float e = expf(sc[j] - mx);
    sc[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  __syncwarp();
  const float inv = 1.f / sum;
  for (int i = lane; i < hd; i += 32) {
    float a = 0.f;
    const T* vcol = v + row0 * ld + hk * hd + i;
    for (int j = 0; j < n_keys; ++j) a = fmaf(sc[j], to_f32(vcol[static_cast<size_t>(j) * ld]), a);
    o[(row0 + qi) * ld_o + h * hd + i] = from_f32<T>(a * inv);
  }
}

template <int HD, bool CAUSAL, bool ROPE>
int launch_flash(const AttnArgs& a, cudaStream_t stream) {
  auto kfn = flash_attn_bf16_kernel<HD, CAUSAL, ROPE>;
  constexpr int SMEM = (BQ + 2 * BKV) * (HD + 8) * 2;
  if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), SMEM)) return rc;
  dim3 grid(ceil_div(a.N, BQ), a.heads_q, a.B);
  const float sl2 = a.scale * 1.4426950408889634f;
  kfn<<<grid, 128, SMEM, stream>>>(
      static_cast<const __nv_bfloat16*>(a.q), static_cast<const __nv_bfloat16*>(a.k),
      static_cast<const __nv_bfloat16*>(a.v), a.ld_qkv, static_cast<__nv_bfloat16*>(a.o), a.ld_o,
      a.N, a.heads_q, a.heads_kv, sl2, a.rope_cos, a.rope_sin);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

template <typename T>
int launch_simt(const AttnArgs& a, cudaStream_t stream) {
  auto kfn = attn_simt_kernel<T>;
  const int smem = 4 * (a.head_dim + a.N) * static_cast<int>(sizeof(float));
  FVLA_REQUIRE(smem <= 200 * 1024, "attention_simt: sequence too long for the score buffer");
  if (smem > 48 * 1024)
    if (int rc = ensure_dyn_smem(reinterpret_cast<const void*>(kfn), smem)) return rc;
  dim3 grid(ceil_div(a.N, 4), a.heads_q, a.B);
  kfn<<<grid, 128, smem, stream>>>(static_cast<const T*>(a.q), static_cast<const T*>(a.k),
                                   static_cast<const T*>(a.v), a.ld_qkv, static_cast<T*>(a.o),
                                   a.ld_o, a.N, a.heads_q, a.heads_kv, a.head_dim, a.scale,
                                   a.causal, a.rope_cos, a.rope_sin);
  FVLA_CUDA_CHECK(cudaGetLastError());
  return 0;
}

int check_attn(const AttnArgs& a) {
  FVLA_REQUIRE(a.B > 0 && a.N > 0 && a.heads_q > 0 && a.heads_kv > 0, "attention: empty problem");
  FVLA_REQUIRE(a.heads_q % a.heads_kv == 0, "attention: heads_q must be a multiple of heads_kv");
  FVLA_REQUIRE(a.head_dim % 8 == 0 && a.ld_qkv % 8 == 0 && a.ld_o % 8 == 0,
               "attention: head_dim / pitches must be multiples of 8");
  FVLA_REQUIRE((a.rope_cos == nullptr) == (a.rope_sin == nullptr), "attention: rope tables");
  return 0;
}

}  // namespace

int attention_simt(int dtype, const AttnArgs& a, cudaStream_t stream) {
  if (int rc = check_attn(a)) return rc;
  if (dtype == DT_F32) return launch_simt<float>(a, stream);
  return launch_simt<__nv_bfloat16>(a, stream);
}

int attention(int dtype, const AttnArgs& a, cudaStream_t stream) {
  if (int rc = check_attn(a)) return rc;
  if (dtype == DT_F32) return launch_simt<float>(a, stream);
  // short causal prefills (T' = 272: three 128-row blocks, the last one nearly empty) keep the 64-query kernel
  if (attention_tc_supported(a)) return attention_tc(a, stream);  // FastViTHD MHSA shapes: tensor-memory flash kernel
  if (attention_v2_supported(a)) return attention_v2(a, stream);
  const bool rope = a.rope_cos != nullptr;
  if (a.head_dim == 32 && !a.causal && !rope) return launch_flash<32, false, false>(a, stream);
  if (a.head_dim == 64 && a.causal && rope) return launch_flash<64, true, true>(a, stream);
  if (a.head_dim == 128 && a.causal && rope) return launch_flash<128, true, true>(a, stream);
  if (a.head_dim == 64 && !a.causal && !rope) return launch_flash<64, false, false>(a, stream);
  if (a.head_dim == 64 && a.causal && !rope) return launch_flash<64, true, false>(a, stream);
  if (a.head_dim == 128 && a.causal && !rope) return launch_flash<128, true, false>(a, stream);
  set_error("attention: no flash instantiation for head_dim=" + std::to_string(a.head_dim) +
            " causal=" + std::to_string(a.causal) + " rope=" + std::to_string(rope ? 1 : 0));
  return 2;
}

}  // namespace fvla
