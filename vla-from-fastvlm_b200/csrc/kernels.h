// Host-callable launchers for every device kernel on the FastVLA forward path.
// All tensors are NHWC / token-major. `dtype` selects the activation + GEMM-weight storage type
// (DT_F32: fp32 parity mode, DT_BF16: throughput mode); accumulation is always fp32.
// Every launcher is stream-ordered, allocates nothing and returns 0 on success (see last_error()).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace fvla {

enum DType : int { DT_F32 = 0, DT_BF16 = 1, DT_U8 = 2 };
inline size_t dtype_size(int dt) { return dt == DT_F32 ? 4 : (dt == DT_BF16 ? 2 : 1); }

// ---- GEMM: D[M,N] = epi(A[M,K] * W[N,K]^T) -----------------------------------------------
struct GemmArgs {
  const void* A = nullptr; int lda = 0;   // [M, K]
  const void* W = nullptr; int ldw = 0;   // [N, K]  (torch Linear / 1x1-conv layout)
  void* D = nullptr; int ldd = 0;         // [M, N]  ([M, N/2] with swiglu)
  int M = 0, N = 0, K = 0;
  const float* bias = nullptr;            // [N]
  const float* row_scale = nullptr;       // [M], multiplies the accumulator (fused RMSNorm rstd)
  const void* resid = nullptr; int ldr = 0;  // [M, N], added after the activation; may alias D
  int act = 0;                            // Act enum
  int swiglu = 0;                         // columns are interleaved (gate, up): out = silu(g) * u
  int block_n = 0;                        // 0 = auto (bf16 path only)
  int ab_f16 = 0;                         // bf16 path: A and W hold FP16 bits (fp16 x fp16 -> fp32 MMA)
  int out_f32 = 0;                        // bf16 path: D is FP32 (ldd in fp32 elements); resid must be null or D itself
                                          // (in-place stream update, done as a TMA reduce-add)
  // bf16 path, Qwen2 qkv projection: rotate-half RoPE (head_dim 64) in the epilogue.  Output columns [0, rope_cols)
  // are heads of 64 columns rotated with table row (m % rope_T); rope_tab [rope_T][32] holds (cos, sin) as packed
  // fp16 pairs (transformers apply_rotary_pos_emb; rope_T >= 32)
  const uint32_t* rope_tab = nullptr; int rope_T = 0, rope_cols = 0;
  // bf16 path, out_f32 with resid == D: split the K loop over `split_k` CTA pairs per tile, every split reduce-adding
  // its partial product into D — for small-M GEMMs whose tiles cannot fill the SMs (the b = 1 prefill)
  int split_k = 0;
};
int gemm_bf16(const GemmArgs& g, cudaStream_t stream);  // tcgen05 + TMA + TMEM
int gemm_f32(const GemmArgs& g, cudaStream_t stream);   // FFMA, fp32 parity mode
inline int gemm(int dtype, const GemmArgs& g, cudaStream_t s) {
  return dtype == DT_BF16 ? gemm_bf16(g, s) : gemm_f32(g, s);
}

// ---- fused ConvFFN tail (bf16, C in {96, 192}): out = resid + W2 . gelu(W1 . x + b1) + b2 ------------------
// W1/b1 are stored PRE-HALVED (the GELU epilogue takes x/2, like ACT_GELU_HALF); the 4x hidden tensor stays on
// chip (ffn_fused_sm100.cu).
struct FfnFusedArgs {
  const void* x = nullptr;      // [M, C]
  const void* w1 = nullptr;     // [hidden, C]  (x 1/2)
  const float* b1 = nullptr;    // [hidden]     (x 1/2)
  const void* w2 = nullptr;     // [C, hidden] as FP16 (the on-chip hidden tensor is fp16)
  const float* b2 = nullptr;    // [C]
  const void* resid = nullptr;  // [M, C], may alias out
  void* out = nullptr;          // [M, C]
  int M = 0, C = 0, hidden = 0;
};
bool ffn_fused_supported(int dtype, int C, int hidden);
int ffn_fused(const FfnFusedArgs& a, cudaStream_t stream);
// C = 384 (stage 2): 64-wide hidden chunks, H double-buffered in tensor memory (ffn_wide_sm100.cu); reached through
// ffn_fused / ffn_fused_supported, opt-in by FVLA_ENABLE_FFN_WIDE
bool ffn_wide_supported(int dtype, int C, int hidden);
int ffn_wide(const FfnFusedArgs& a, cudaStream_t stream);

// ---- image ingest ----------------------------------------------------------------------------
struct PreprocessArgs {
  const void* src = nullptr; int src_dtype = DT_F32;  // DT_F32 / DT_U8 / DT_BF16
  int src_nhwc = 0;         // 0: (B,C,h,w)   1: (B,h,w,C)
  int B = 0, C = 3, h = 0, w = 0;
  int S = 0;                // square output side
  int letterbox = 1;        // aspect-preserving resize + left/top pad, else plain stretch
  float pad_value = 0.f;
  float scale = 1.f;        // value scale applied before the optional mean/std
  int normalize = 0; float mean[3] = {0, 0, 0}; float inv_std[3] = {1, 1, 1};
  void* dst = nullptr;      // (B,S,S,3) in `dtype`
};
int preprocess_images(int dtype, const PreprocessArgs& a, cudaStream_t stream);

// ---- convolutions (NHWC) ---------------------------------------------------------------------
// Dense 3x3 stride-2 pad-1 conv with 3 input channels (FastViTHD stem.0) + bias + GELU.
// w_packed: [27][Cout] fp32 with row = (ky*3+kx)*3 + ci.
int stem_conv3x3_s2(int dtype, const void* in, const float* w_packed, const float* bias, void* out,
                    int B, int H, int W, int Cout, cudaStream_t stream);
// bf16 tensor-core route for the same conv: gather 27-tap patches into [B*(H/2)*(W/2), 32] (cols 27..31 zero)
int stem_im2col_bf16(const void* in, void* col, int B, int H, int W, cudaStream_t stream);
// stem.0 + stem.1 fused (bf16): [B,S,S,4] -> GELU(dw3x3 s2(GELU(conv3x3 s2 + b0)) + b1) -> [B,S/4,S/4,C]; the
// S/2-sided intermediate map stays in shared memory (stem_fused.cu).  btab from stem_fused_build_btab (host).
bool stem_fused_supported(int dtype, int S, int C);
size_t stem_fused_btab_words(int C);
void stem_fused_build_btab(const float* w0_packed, int C, uint32_t* btab_host);
int stem_fused(const void* in, const uint32_t* btab, const float* b0_half, const float* w1_packed, const float* b1,
               void* out, int B, int S, int C, cudaStream_t stream);
// Depthwise / grouped k x k conv, groups = Cin, Cout = mult*Cin (mult 1 or 2), pad k/2,
// + bias (+ GELU).  w_packed: [k*k][Cout] fp32.
// `wtab` (optional): the tensor-core 7x7 path's Toeplitz table from dwconv7_mma_prepare(); built on the
// fly when null.
int dwconv(int dtype, const void* in, const float* w_packed, const float* bias, void* out, int B,
           int H, int W, int Cin, int mult, int ksize, int stride, int act, cudaStream_t stream,
           const uint32_t* wtab = nullptr);

// smem-tiled stride-2 x2-channel grouped 7x7 (patch-embed proj.0), bf16, Wo%32==0, Ho%8==0, Cout%32==0
bool dwconv_s2m2_tiled_supported(int dtype, int H, int W, int Cin, int mult, int k, int stride);
int dwconv_s2m2_tiled(const void* in, const float* w_packed, const float* bias, void* out, int B, int H, int W,
                      int Cin, int act, cudaStream_t stream);

// tensor-core depthwise 7x7 (bf16, stride 1, no activation, C%32==0, H%16==0, W%32==0 or W==16): dwconv_mma.cu
bool dwconv7_mma_supported(int dtype, int H, int W, int C, int mult, int k, int stride, int act);
size_t dwconv7_wtab_bytes(int C);
int dwconv7_mma_prepare(const float* w_packed, int C, uint32_t* wtab, cudaStream_t stream);
// 7x7 stride 2 with two output channels per input channel (+ optional GELU) on the same tensor-core pipeline; its
// table (same size) comes from dwconv7_s2m2_mma_prepare
bool dwconv7_s2m2_mma_supported(int dtype, int H, int W, int C, int mult, int k, int stride, int act);
int dwconv7_s2m2_mma_prepare(const float* w_packed, int C, uint32_t* wtab, cudaStream_t stream);
int dwconv7_s2m2_mma(const void* in, const uint32_t* wtab, const float* bias, void* out, int B, int H, int W, int C,
                     int act, cudaStream_t stream);
int dwconv7_mma(const void* in, const uint32_t* wtab, const float* bias, void* out, int B, int H, int W,
                int C, cudaStream_t stream);

// persistent TMA-fed depthwise 3x3 (bf16, stride 1, no activation, C%32==0, H%8==0, W%32==0), fp32 FFMA2 taps: dwconv3_tma.cu
bool dwconv3_tma_supported(int dtype, int H, int W, int C, int mult, int k, int stride, int act);
int dwconv3_tma(const void* in, const float* w_packed, const float* bias, void* out, int B, int H, int W, int C,
                cudaStream_t stream);

// smem-tiled bf16 fast path (stride 1, mult 1, k in {3,7}, W%64==0, H%8==0, C%32==0); dwconv() uses it
bool dwconv_tiled_supported(int dtype, int H, int W, int C, int mult, int k, int stride);
int dwconv_tiled(const void* in, const float* w_packed, const float* bias, void* out, int B, int H,
                 int W, int C, int k, int act, cudaStream_t stream);

// ---- squeeze-excite tail of conv_exp ---------------------------------------------------------
// x: [B, HW, C]; gate = sigmoid(W2 relu(W1 mean_hw(x) + b1) + b2); out = gelu(x * gate)
int se_gelu(int dtype, const void* x, void* out, int B, int HW, int C, int Cr, const float* w1,
            const float* b1, const float* w2, const float* b2, float* scratch_mean,
            float* scratch_gate, cudaStream_t stream);

// ---- attention -------------------------------------------------------------------------------
struct AttnArgs {
  const void* q = nullptr; const void* k = nullptr; const void* v = nullptr;
  int ld_qkv = 0;           // row pitch (elements) shared by q/k/v
  void* o = nullptr; int ld_o = 0;
  int B = 0, N = 0;         // N tokens per sample (rows of sample b start at b*N)
  int heads_q = 0, heads_kv = 0, head_dim = 0;
  float scale = 1.f;
  int causal = 0;
  const float* rope_cos = nullptr;  // [>=N][head_dim/2], null = no rotary embedding
  const float* rope_sin = nullptr;
};
int attention(int dtype, const AttnArgs& a, cudaStream_t stream);       // bf16: mma.sync flash; f32: SIMT
int attention_simt(int dtype, const AttnArgs& a, cudaStream_t stream);
// bf16, head_dim 32, non-causal, N % 128 == 0 (FastViTHD MHSA): tcgen05 MMAs with S/P/O in tensor memory (attention_sm100.cu)
bool attention_tc_supported(const AttnArgs& a);
int attention_tc(const AttnArgs& a, cudaStream_t stream);
// bf16, no rotary: 128-query CTAs, cp.async double buffering, ldmatrix operands (attention_v2.cu)
bool attention_v2_supported(const AttnArgs& a);
int attention_v2(const AttnArgs& a, cudaStream_t stream);  // reference-grade SIMT for either dtype

// ---- Qwen2 glue ------------------------------------------------------------------------------
// x_f32 (bf16 mode only): x is the decoder's FP32 residual stream, out stays bf16
int rmsnorm(int dtype, const void* x, const float* weight, void* out, int rows, int H, float eps,
            cudaStream_t stream, int x_f32 = 0);
// LayerNorm over the channel dimension of token-major rows WITHOUT the affine part (FastViTHD `LayerNormChannel`
// [EXT mci.py]; weight/bias are folded into the GEMM that follows): out = (x - mean) * rsqrt(var + eps), C % 8 == 0
int layernorm_rows(int dtype, const void* x, void* out, int rows, int C, float eps, cudaStream_t stream);
// plan[b*T+t]: >=0 vocab id, -1 zero row (padding), <=-2 image row (-2-idx) of sample b
// out_f32 (bf16 mode only): write the rows as FP32 (the decoder's residual stream)
int embed_splice(int dtype, const void* table, const void* img_feats, int n_img, const int* plan,
                 void* out, int B, int T, int H, cudaStream_t stream, int out_f32 = 0);
// mode 0: last_token at pool_idx[b]; mode 1: mean over t < len[b]. Applies the final RMSNorm.
int pool_norm(int dtype, const void* hidden, const float* norm_w, const int* pool_idx,
              const int* lens, int mode, float* pooled, int B, int T, int H, float eps,
              cudaStream_t stream);
// rotate-half RoPE in place on the first `heads` head slices (q heads then k heads) of the qkv rows
int rope_inplace(int dtype, void* qkv, int ld, int B, int n_tok, int heads, int head_dim,
                 const float* cos_t, const float* sin_t, cudaStream_t s);
int swiglu_interleaved(int dtype, const void* gu, void* out, int rows, int I, cudaStream_t stream);

// ---- FastVLA action head (fastvla/fastvlm_with_expert.py:23-38, 50-54) ------------------------
struct HeadWeights {  // all fp32 except the four matrices, which are in `dtype`
  const float* ln_s_w; const float* ln_s_b;        // state_projection.0
  const void* w_state; const float* b_state;       // state_projection.1  [Hd, S]
  const void* w_f0; const float* b_f0;             // fusion.0            [F, H+Hd]
  const float* ln_f_w; const float* ln_f_b;        // fusion.1
  const void* w_f4; const float* b_f4;             // fusion.4            [F, F]
  const void* w_act; const float* b_act;           // action_head         [A, F]
  // LeRobot (un)normaliser steps fused at both ends of the head (lerobot_fastvla/processor_fastvla.py:30-48), device
  // fp32, never null (identity by default): state' = (state - st_mean) * st_inv_std;  action' = action * act_scale + act_shift
  const float* st_mean; const float* st_inv_std;   // [S]
  const float* act_scale; const float* act_shift;  // [A]
  int H, S, Hd, F, A;
};
// pooled [B,H] fp32, states [B,S] fp32 -> actions [B,A] fp32.  state_feat [B,Hd], x1_scratch [B,F], fused [B,F]
// (fp32) are the rows the kernel's CTA cluster exchanges layer outputs through; state_feat / fused are also the
// parity taps.
int action_head(int dtype, const HeadWeights& w, const float* pooled, const float* states,
                float* actions, float* state_feat, float* x1_scratch, float* fused, int B,
                cudaStream_t stream);

// ---- FastVLA head training step (fp32): forward + MSE + backward, gradients into a flat buffer (head_train.cu) ----
struct HeadTrainArgs {
  int B = 0, H = 0, S = 0, Hd = 0, F = 0, A = 0;
  HeadWeights w{};                      // all twelve tensors as DEVICE FP32 (the torch parameters themselves)
  const float* pooled = nullptr;        // [B, H] backbone features (no gradient: the backbone is frozen)
  const float* states = nullptr;        // [B, S]
  const float* target = nullptr;        // [B, A]
  const uint8_t* keep_mask = nullptr;   // [B, F] Dropout keep mask (1 = keep) when drop_p > 0
  float drop_p = 0.f;
  float* grads = nullptr;               // flat, nn.Module parameter order (written, not accumulated)
  float* loss = nullptr;                // device scalar
  float* actions = nullptr;             // optional [B, A] predictions
  float* scratch = nullptr; size_t scratch_floats = 0;
};
size_t head_train_scratch_floats(int B, int H, int S, int Hd, int F, int A);
int head_train_step(const HeadTrainArgs& t, cudaStream_t stream);

// ---- small utilities -------------------------------------------------------------------------
int convert(int src_dtype, const void* src, int dst_dtype, void* dst, long long n, cudaStream_t s);
int rope_table(float* cos_t, float* sin_t, int T, int head_dim, float theta, cudaStream_t s);

}  // namespace fvla
