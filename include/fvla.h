/*
 * fvla.h — C ABI of the B200-native FastVLA policy forward (libfvla.so).
 *
 * Drop-in boundary for ONE path of syun88/VLA-from-FastVLM: observation images + prompt token ids
 * + robot state -> action, i.e. what `FastVLMWithExpert.forward` computes
 * (src/vla_fastvlm/fastvla/fastvlm_with_expert.py:40-54) through `FastVLMBackbone.forward`
 * (src/vla_fastvlm/model/fastvlm_adapter.py:501-560) and the remote-code
 * `LlavaQwen2ForCausalLM` it calls at fastvlm_adapter.py:533 (FastViTHD + mlp2x_gelu projector +
 * LLaVA image splice + Qwen2 prefill).
 *
 * The reference has no FFI of its own (it is pure Python dispatching to ATen/cuDNN/cuBLAS); every
 * entry point below therefore cites the Python call site whose arithmetic it replaces.  Plain
 * pointers and sizes only, no torch types.  Conventions:
 *   - every function returns 0 on success; on failure fvla_last_error() describes it (thread-local)
 *   - "device" pointers are CUDA device pointers of the current device; "host" pointers are CPU memory
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no hidden device syncs
 *   - the caller owns all buffers it passes; the engine owns its packed weights and workspace
 *   - a handle is bound to one device and is not thread-safe (one policy per process, as in the reference)
 *   - there is NO CPU path: without a CUDA device every compute call fails.
 */
#ifndef FVLA_H_
#define FVLA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVLA_ABI_VERSION 1
#define FVLA_IMAGE_TOKEN_INDEX (-200) /* LLaVA placeholder id spliced by the backbone [EXT] */

enum { FVLA_F32 = 0, FVLA_BF16 = 1, FVLA_U8 = 2 };
/* FVLA_ACT_GELU_HALF: operand is x/2 (weights + bias pre-halved by the caller), result gelu(x) */
enum { FVLA_ACT_NONE = 0, FVLA_ACT_GELU = 1, FVLA_ACT_SILU = 2, FVLA_ACT_RELU = 3, FVLA_ACT_GELU_HALF = 4 };
enum { FVLA_POOL_LAST_TOKEN = 0, FVLA_POOL_MEAN = 1 }; /* fastvlm_adapter.py:337-359 */

#define FVLA_MAX_VIS_STAGES 8

/* Architecture + numerics description.  Vision fields follow FastViTHD, language
 * fields the checkpoint's Qwen2 config.json, head fields FastVLAConfig
 * (fastvla/configuration_fastvla.py:9-32). */
typedef struct fvla_config {
  int32_t dtype;                /* FVLA_F32 (parity mode) or FVLA_BF16 (throughput mode) */
  /* FastViTHD */
  int32_t image_size;           /* square side fed to the tower (expected_size, adapter :143) */
  int32_t vis_num_stages;
  int32_t vis_layers[FVLA_MAX_VIS_STAGES];
  int32_t vis_dims[FVLA_MAX_VIS_STAGES];
  int32_t vis_attention[FVLA_MAX_VIS_STAGES]; /* 0 RepMixer token mixer, 1 MHSA */
  int32_t vis_pos_emb[FVLA_MAX_VIS_STAGES];   /* 1: RepCPE 7x7 before the stage */
  int32_t vis_mlp_ratio;        /* 4 */
  int32_t vis_head_dim;         /* 32 */
  int32_t vis_se_reduced;       /* conv_exp SE bottleneck channels (3072/16 = 192) */
  /* Qwen2 */
  int32_t hidden, n_layers, n_q_heads, n_kv_heads, head_dim, intermediate, vocab;
  float rms_eps, rope_theta;
  /* FastVLA head */
  int32_t state_dim, action_dim, hidden_dim, fusion_dim;
  int32_t pool_mode;            /* FVLA_POOL_* */
  /* engine knobs */
  int32_t vision_chunk;         /* images per FastViTHD pass (bounds workspace, keeps tiles in L2); 0 = auto */
  int32_t skip_unused_vision;   /* 1: skip the tower when no row holds an image placeholder (F4) */
} fvla_config;

typedef struct fvla_engine fvla_engine;

/* Forward inputs.  Mirrors FastVLMWithExpert.forward(images, states, tasks) after tokenisation. */
typedef struct fvla_forward_args {
  int32_t batch;
  /* observation images, device memory, any size; resized+letterboxed on the GPU
     (replaces the CPU bounce at fastvlm_adapter.py:479-488) */
  const void* images;
  int32_t img_dtype;            /* FVLA_F32 / FVLA_BF16 / FVLA_U8 */
  int32_t img_nhwc;             /* 0: (B,C,h,w), 1: (B,h,w,C) */
  int32_t img_c, img_h, img_w;
  int32_t letterbox;            /* resize_with_padding (adapter :451-461) */
  float pad_value;
  float img_scale;              /* multiplies pixel values (1/255 path of adapter :472-473) */
  int32_t normalize;            /* ImageNet mean/std (adapter :463-477), default off */
  float mean[3], inv_std[3];
  /* prompt: HOST int32 [batch, n_tokens] right-padded, FVLA_IMAGE_TOKEN_INDEX marks the image slot */
  const int32_t* token_ids;
  const int32_t* text_len;      /* HOST [batch]: attention_mask.sum(1) (adapter :354) */
  int32_t n_tokens;
  const int32_t* pool_idx;      /* HOST [batch] position pooled from the merged sequence, or NULL
                                   for the reference-literal text_len-1 (adapter :353-358) */
  const float* states;          /* device fp32 [batch, state_dim]; NULL = backbone only */
  float* actions;               /* device fp32 [batch, action_dim]; may be NULL with states */
  float* pooled;                /* optional device fp32 [batch, hidden] backbone features */
} fvla_forward_args;

/* ---- engine life cycle ---------------------------------------------------------------------- */
int fvla_abi_version(void);
const char* fvla_last_error(void);
int fvla_device_count(int* count);

int fvla_create(const fvla_config* cfg, fvla_engine** out);
void fvla_destroy(fvla_engine* e);

/* Stage one tensor under its FastVLMWithExpert.state_dict() key, e.g.
 *   "backbone.model.model.vision_tower.vision_tower.model.network.0.0.convffn.fc1.weight"
 *   "backbone.model.model.layers.3.self_attn.q_proj.bias", "fusion.0.weight", "action_head.bias".
 * `data` is HOST memory (fp32 or bf16) and is copied.  Replaces load_state_dict on the reference
 * modules (utils/checkpoint.py:14-47). */
int fvla_load_tensor(fvla_engine* e, const char* name, const void* data, int32_t dtype,
                     int32_t ndim, const int64_t* shape);
/* Number of tensors fvla_finalize still needs; names are written '\n'-separated into buf. */
int fvla_missing_tensors(fvla_engine* e, char* buf, int64_t buf_len, int32_t* n_missing);
/* Fold BatchNorm / layer-scale, fuse q|k|v and gate|up, repack for the kernels, upload. */
int fvla_finalize(fvla_engine* e);

/* Pre-size the workspace (otherwise grown on first use; growth is not stream-capture safe). */
int fvla_reserve(fvla_engine* e, int32_t batch, int32_t n_tokens);
int64_t fvla_workspace_bytes(fvla_engine* e);
int64_t fvla_weight_bytes(fvla_engine* e);

/* The hot path: FastVLMWithExpert.forward (fastvlm_with_expert.py:40-54). */
int fvla_forward(fvla_engine* e, const fvla_forward_args* args, void* stream);

/* LeRobot's NormalizerProcessorStep (STATE, mean/std) and UnnormalizerProcessorStep (ACTION) run as separate
 * elementwise passes around the policy (lerobot_fastvla/processor_fastvla.py:30-48).  Here they are the first and last
 * arithmetic of the action-head kernel:  state' = (state - state_mean) * state_inv_std;
 * action' = action * action_scale + action_shift.  HOST fp32 vectors ([state_dim], [state_dim], [action_dim],
 * [action_dim]); any may be NULL (identity for that part).  Takes effect from the next fvla_forward; call after
 * fvla_finalize.  With all four NULL the head computes exactly FastVLMWithExpert.forward. */
int fvla_set_io_normalization(fvla_engine* e, const float* state_mean, const float* state_inv_std,
                              const float* action_scale, const float* action_shift);

/* Number of kernels the last fvla_forward launched, and algorithmic FLOPs (SURVEY §8d formula). */
int64_t fvla_last_launch_count(fvla_engine* e);
double fvla_last_forward_flops(fvla_engine* e);

/* Per-stage parity taps: before a forward, register a device buffer for a stage id; the forward
 * copies that stage's tensor into it (engine dtype, NHWC / token-major).  dst NULL unregisters. */
enum {
  FVLA_TAP_PREPROCESS = 0,  /* [B,S,S,4] (4th channel zero) */
  FVLA_TAP_STEM = 1,        /* [B,S/4,S/4,d0] */
  FVLA_TAP_VIS_STAGE0 = 10, /* +i: output of FastViTHD stage i, [B,h,w,d_i] */
  FVLA_TAP_IMAGE_FEATURES = 30, /* conv_exp output [B,n_img,mm_hidden] */
  FVLA_TAP_PROJECTOR = 31,  /* [B,n_img,H] */
  FVLA_TAP_EMBEDS = 32,     /* inputs_embeds after the splice [B,T',H] */
  FVLA_TAP_LAYER0 = 100,    /* +l: decoder layer l output [B,T',H] */
  FVLA_TAP_POOLED = 1000,   /* fp32 [B,H] */
  FVLA_TAP_STATE_FEAT = 1001, /* fp32 [B,hidden_dim] */
  FVLA_TAP_FUSED = 1002     /* fp32 [B,fusion_dim] */
};
int fvla_set_tap(fvla_engine* e, int32_t stage, void* dst, int64_t capacity_bytes);
int fvla_merged_len(fvla_engine* e);  /* T' of the last forward */

/* Per-launch timing with CUDA events on the forward's stream (off by default).  The report is CSV:
 * label,count,total_ms,flops,bytes — one line per distinct kernel/shape, device-synchronising. */
int fvla_set_profile(fvla_engine* e, int32_t on);
int fvla_profile_report(fvla_engine* e, char* buf, int64_t buf_len);

/* ---- single-kernel entry points (used by the per-kernel parity tests and micro-benchmarks) ----
 * All pointers are device pointers; dtype selects fp32 / bf16 activations. */

/* nn.Linear / 1x1 Conv2d + bias + activation + residual [EXT FastViTHD ConvFFN, mm_projector;
 * transformers Qwen2MLP / Qwen2Attention projections]: D = act(rs*A W^T + bias) + resid.
 * `swiglu` is a flag word: bit 0 = SwiGLU epilogue over interleaved (gate, up) columns; bit 1 (bf16 path) = A and
 * W hold FP16 bits; bit 2 (bf16 path) = D is FP32 (the decoder's residual stream, ldd in fp32 elements) and resid,
 * if given, must be D itself (in-place update); bits 8..15 = split-K factor of that in-place update (0/1 = none).  act 5 (bf16 path) = GELU of a pre-halved operand stored as FP16 (ConvFFN hidden tensor). */
int fvla_op_gemm(int32_t dtype, const void* A, int32_t lda, const void* W, int32_t ldw, void* D,
                 int32_t ldd, int32_t M, int32_t N, int32_t K, const float* bias,
                 const float* row_scale, const void* resid, int32_t ldr, int32_t act,
                 int32_t swiglu, int32_t block_n, void* stream);
/* fastvlm_adapter.py:479-488 image canonicalisation; dst [B,S,S,4] */
int fvla_op_preprocess(int32_t dtype, const void* src, int32_t src_dtype, int32_t nhwc, int32_t B,
                       int32_t C, int32_t h, int32_t w, int32_t S, int32_t letterbox,
                       float pad_value, float scale, int32_t normalize, const float* mean3,
                       const float* inv_std3, void* dst, void* stream);
/* FastViTHD stem.0: Conv2d(3,C,3,stride 2,pad 1)+GELU on [B,H,W,4] -> [B,H/2,W/2,C]; w [27][C] */
int fvla_op_stem_conv(int32_t dtype, const void* in, const float* w_packed, const float* bias,
                      void* out, int32_t B, int32_t H, int32_t W, int32_t Cout, void* stream);
/* depthwise/grouped Conv2d(groups=Cin) k in {3,7}, stride {1,2}, mult {1,2}; w [k*k][Cout] */
int fvla_op_dwconv(int32_t dtype, const void* in, const float* w_packed, const float* bias,
                   void* out, int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t mult,
                   int32_t ksize, int32_t stride, int32_t act, void* stream);
/* conv_exp squeeze-excite + GELU; scratch_mean/gate: fp32 [B,C] */
int fvla_op_se_gelu(int32_t dtype, const void* x, void* out, int32_t B, int32_t HW, int32_t C,
                    int32_t Cr, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* scratch_mean, float* scratch_gate, void* stream);
/* MHSA / Qwen2 GQA attention on slices of the fused qkv buffer; impl 0 = production, 1 = SIMT */
int fvla_op_attention(int32_t dtype, int32_t impl, const void* q, const void* k, const void* v,
                      int32_t ld_qkv, void* o, int32_t ld_o, int32_t B, int32_t N,
                      int32_t heads_q, int32_t heads_kv, int32_t head_dim, float scale,
                      int32_t causal, const float* rope_cos, const float* rope_sin, void* stream);
int fvla_op_rmsnorm(int32_t dtype, const void* x, const float* weight, void* out, int32_t rows,
                    int32_t H, float eps, void* stream);
/* FastViTHD `LayerNormChannel` [EXT mci.py] without its affine part (folded into the qkv GEMM by fvla_finalize):
 * per row of x [rows, C]: out = (x - mean) * rsqrt(var + eps), statistics in fp32 */
int fvla_op_layernorm_rows(int32_t dtype, const void* x, void* out, int32_t rows, int32_t C, float eps,
                           void* stream);
/* FastViTHD ConvFFN tail fused on chip (bf16, C in {96,192}, hidden %128 == 0) [EXT convffn.fc1 -> GELU -> fc2,
 * + layer-scaled residual]: out = resid + w2 . gelu(w1 . x + b1) + b2.  w1 [hidden,C] and b1 are passed
 * PRE-HALVED (the epilogue evaluates gelu from x/2); w2 [C,hidden]; resid may alias out. */
int fvla_op_ffn_fused(const void* x, const void* w1_half, const float* b1_half, const void* w2,
                      const float* b2, const void* resid, void* out, int32_t M, int32_t C,
                      int32_t hidden, void* stream);
/* ---- training step of the action head (BASELINE config 3) ----------------------------------------------------
 * What `loss.backward()` does to the head in the reference's step (training/trainer.py:168-182: compute_loss ->
 * accelerator.backward; loss = F.mse_loss(pred, gt[:, 0]), lerobot_fastvla/modeling_fastvla.py:127-133; the backbone
 * runs under no_grad, fastvlm_adapter.py:501): forward of the head modules of fastvla/fastvlm_with_expert.py:23-38
 * in TRAIN mode (Dropout active), MSE loss, backward — in fp32, on the torch parameters themselves.
 *   params[12]: DEVICE fp32 tensors in nn.Module parameter order — state_projection.0.{weight,bias},
 *               state_projection.1.{weight,bias}, fusion.0.{weight,bias}, fusion.1.{weight,bias},
 *               fusion.4.{weight,bias}, action_head.{weight,bias}
 *   pooled [B,H] backbone features (fvla_forward's `pooled` output), states [B,S], target [B,A]: device fp32
 *   keep_mask [B,F] uint8 (1 = keep) when drop_p > 0, else NULL
 *   grads: the flat fp32 gradient buffer, same order and sizes as params (WRITTEN: this is d loss / d param, to be
 *          all-reduced over the data-parallel ranks by the caller — it is the buffer NCCL runs over, no copy)
 *   loss: device scalar;  actions: optional [B,A] predictions;  scratch: fvla_head_train_scratch_floats() floats */
int64_t fvla_head_train_scratch_floats(int32_t B, int32_t H, int32_t S, int32_t Hd, int32_t F, int32_t A);
int fvla_head_forward_backward(int32_t B, int32_t H, int32_t S, int32_t Hd, int32_t F, int32_t A,
                               const float* const* params, const float* pooled, const float* states,
                               const float* target, const uint8_t* keep_mask, float drop_p, float* grads, float* loss,
                               float* actions, float* scratch, int64_t scratch_floats, void* stream);
int fvla_op_convert(int32_t src_dtype, const void* src, int32_t dst_dtype, void* dst, int64_t n,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FVLA_H_ */
