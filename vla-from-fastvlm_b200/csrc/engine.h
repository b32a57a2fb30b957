// FastVLA forward engine: owns packed weights + workspace and sequences the sm_100a kernels for
// FastViTHD -> mm_projector -> LLaVA splice -> Qwen2 prefill -> pooling -> action head.
#pragma once
#include "../../include/fvla.h"
#include "kernels.h"

#include <map>
#include <string>
#include <vector>

namespace fvla {

struct HostTensor {
  std::vector<float> data;
  std::vector<int64_t> shape;
  int64_t numel() const { int64_t n = 1; for (auto d : shape) n *= d; return n; }
};

struct GemmW {  // packed nn.Linear / 1x1 conv
  void* w = nullptr;          // [N, K] in engine dtype
  void* w_f16 = nullptr;      // same matrix as fp16: ConvFFN fc2 (bf16 mode), whose hidden operand is stored as fp16
  const float* bias = nullptr;  // [N] fp32 or null
  int N = 0, K = 0;
  bool half_in = false;       // weights+bias pre-scaled by 1/2: the GELU epilogue takes x/2 (ACT_GELU_HALF)
};
struct DwW {  // packed depthwise / grouped conv
  float* w = nullptr;     // [k*k][Cout] fp32
  float* bias = nullptr;  // [Cout] fp32
  uint32_t* wtab = nullptr;  // Toeplitz B-fragment table of the tensor-core 7x7 path (bf16 mode)
  int cin = 0, mult = 1, k = 3, stride = 1, act = 0;
};
struct VisBlock {
  bool attn = false;
  bool attn_layernorm = false;  // pre-attention norm is LayerNormChannel (run-time statistics) instead of BatchNorm2d
  DwW mixer;           // RepMixer reparam 3x3 (repmixer blocks only)
  GemmW qkv, proj;     // attention blocks only (BN folded into qkv, layer_scale_1 into proj)
  DwW ffn_dw;          // ConvFFN 7x7 with BN folded
  GemmW fc1, fc2;      // fc2 carries the folded layer scale
};
struct VisStage {
  int dim = 0;
  bool has_cpe = false; DwW cpe;
  std::vector<VisBlock> blocks;
  bool has_down = false; DwW down_dw; GemmW down_pw;
};
struct DecLayer {
  float* ln1 = nullptr; float* ln2 = nullptr;
  GemmW qkv, o, gate_up, down;
};

struct Workspace {
  std::map<std::string, std::pair<void*, size_t>> bufs;
  size_t total = 0;
};

class Engine {
 public:
  explicit Engine(const fvla_config& cfg);
  ~Engine();
  int load_tensor(const char* name, const void* data, int dtype, int ndim, const int64_t* shape);
  int missing(std::vector<std::string>* out);
  int finalize();
  int reserve(int B, int n_tokens);
  int forward(const fvla_forward_args& a, cudaStream_t stream);
  int set_tap(int stage, void* dst, int64_t cap);
  // host fp32 vectors ([state_dim] x2, [action_dim] x2), any may be null (= identity for that part)
  int set_io_normalization(const float* state_mean, const float* state_inv_std, const float* action_scale,
                           const float* action_shift);
  // per-launch CUDA-event profiling (off by default; adds two event records per kernel)
  void set_profile(bool on) { profile_ = on; }
  int profile_report(std::string* csv);  // syncs, aggregates by label, clears the samples

  fvla_config cfg;
  int64_t launches = 0;
  double flops = 0.0;
  int merged_len = 0;
  size_t weight_bytes = 0;
  size_t workspace_bytes() const { return ws_.total; }

 private:
  std::vector<std::string> required_names() const;
  const HostTensor* find(const std::string& name) const;
  int need(const std::string& name, const HostTensor** out, std::initializer_list<int64_t> shape);

  float* upload_f32(const std::vector<float>& v);
  void* upload_act(const std::vector<float>& v);  // engine dtype
  int make_gemm(GemmW* g, const std::vector<float>& w, int N, int K, const std::vector<float>* bias,
                bool gelu_half = false);
  int make_dw(DwW* d, const std::vector<float>& w_oihw, const std::vector<float>& bias, int cin,
              int mult, int k, int stride, int act);
  int pack_vision();
  int pack_decoder();
  int pack_head();
  int update_head_tensor(const std::string& name, const HostTensor& t);
  struct HeadSlot { void* ptr; int64_t n; bool act; };
  std::map<std::string, HeadSlot> head_slots_;

  int ensure(const std::string& name, size_t bytes, void** out);
  int tap(int stage, const void* src, size_t bytes, size_t dst_offset_bytes, cudaStream_t s);

  // rope: rotate the q and k heads in the epilogue (qkv projection, head_dim 64); split_k: see GemmArgs
  int run_gemm(const GemmW& w, const void* A, void* D, int M, int act, const void* resid, bool swiglu,
               cudaStream_t s, bool ab_f16 = false, bool out_f32 = false, bool rope = false, int split_k = 0,
               int block_n = 0);
  int run_ffn(const VisBlock& blk, const void* z, void* hid, void* out_resid, int M, cudaStream_t s);
  int run_dw(const DwW& w, const void* in, void* out, int B, int H, int W, cudaStream_t s);
  int vision_chunk(const fvla_forward_args& a, int c0, int bc, void* feats, cudaStream_t s);
  int launch_all(const fvla_forward_args& a, int B, int Tm, bool any_image, cudaStream_t s);
  int forward_graph(const fvla_forward_args& a, int B, int Tm, bool any_image, cudaStream_t s);
  void clear_graphs();
  static constexpr int kGraphMaxBatch = 8;  // larger batches are GPU-bound; their inputs are not worth re-staging
  struct GraphEntry { cudaGraphExec_t exec = nullptr; int state = 0; int64_t launches = 0; double flops = 0.0; };
  std::map<std::string, GraphEntry> graphs_;
  cudaStream_t graph_stream_ = nullptr;

  size_t esz() const { return dtype_size(cfg.dtype); }
  int n_img_tokens() const;
  int mm_hidden() const;

  std::map<std::string, HostTensor> host_;
  bool finalized_ = false;
  std::vector<void*> dev_allocs_;

  // packed model
  float* stem0_w_ = nullptr; float* stem0_b_ = nullptr;
  uint32_t* stem_btab_ = nullptr; float* stem0_bh_ = nullptr;  // fused stem: B-fragment table, halved bias
  GemmW stem0_gemm_;  // bf16 mode: im2col form [d0][32]
  DwW stem1_; GemmW stem2_;
  std::vector<VisStage> stages_;
  DwW exp_dw_;
  float *se_w1_ = nullptr, *se_b1_ = nullptr, *se_w2_ = nullptr, *se_b2_ = nullptr;
  GemmW proj0_, proj2_;
  void* embed_ = nullptr;
  std::vector<DecLayer> layers_;
  float* final_norm_ = nullptr;
  float* rope_cos_ = nullptr; float* rope_sin_ = nullptr; int rope_len_ = 0;
  uint32_t* rope_tab_ = nullptr;   // the same table as (cos, sin) fp16 pairs [pos][head_dim/2]: RoPE epilogue of the qkv GEMM
  int merged_T_ = 1;               // T' of the forward being launched (row -> position)
  // algorithmic-work accounting of the decoder: valid rows / padded rows (and the same for squared lengths), so that
  // `flops` counts the tokens of the ragged batch, not the B x T'max rows the GEMMs execute
  double dec_row_frac_ = 1.0, dec_sq_frac_ = 1.0, flop_scale_ = 1.0;
  HeadWeights head_{};
  float* io_norm_ = nullptr;  // [S mean | S inv_std | A scale | A shift]

  struct HostStage { int* p = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; };
  static constexpr int kHostRing = 4;
  HostStage host_ring_[kHostRing];
  int host_ring_next_ = 0;

  Workspace ws_;
  std::map<int, std::pair<void*, int64_t>> taps_;

  struct ProfSample { std::string label; cudaEvent_t e0, e1; double flops, bytes; };
  bool profile_ = false;
  std::vector<ProfSample> prof_;
  std::vector<cudaEvent_t> ev_pool_;
  cudaEvent_t get_event();
  void prof_begin(cudaStream_t s);
  void prof_end(const std::string& label, double fl, double by, cudaStream_t s);
  cudaEvent_t prof_e0_ = nullptr;
  const char* prof_scope_ = "vis.";
};

}  // namespace fvla
