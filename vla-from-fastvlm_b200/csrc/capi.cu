// extern "C" surface of libfvla.so (declared in include/fvla.h).
#include "engine.h"
#include "common.cuh"

#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <utility>

namespace fvla {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
const char* last_error() { return g_error.c_str(); }

bool pdl_enabled() {
  static const bool on = std::getenv("FVLA_DISABLE_PDL") == nullptr;
  return on;
}

int num_sms() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      cached = 148;
  }
  return cached;
}

int ensure_dyn_smem(const void* fn, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<const void*, int>, int> raised;
  int dev = 0;
  FVLA_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_pair(fn, dev);
  auto it = raised.find(key);
  if (it != raised.end() && it->second >= bytes) return 0;
  FVLA_CUDA_CHECK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  raised[key] = bytes;
  return 0;
}

}  // namespace fvla

struct fvla_engine {
  fvla::Engine impl;
  explicit fvla_engine(const fvla_config& c) : impl(c) {}
};

using fvla::set_error;

extern "C" {

int fvla_abi_version(void) { return FVLA_ABI_VERSION; }
const char* fvla_last_error(void) { return fvla::last_error(); }

int fvla_device_count(int* count) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    cudaGetLastError();
    set_error(std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (count) *count = 0;
    return 1;
  }
  if (count) *count = n;
  return 0;
}

int fvla_create(const fvla_config* cfg, fvla_engine** out) {
  if (cfg == nullptr || out == nullptr) { set_error("fvla_create: null argument"); return 2; }
  int n = 0;
  if (fvla_device_count(&n) != 0 || n == 0) {
    set_error("fvla_create: no CUDA device — this library has no CPU path");
    return 1;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("fvla_create: kernels are built for sm_100a only, device has compute capability " +
              std::to_string(major) + ".x");
    return 1;
  }
  *out = new (std::nothrow) fvla_engine(*cfg);
  if (*out == nullptr) { set_error("fvla_create: out of host memory"); return 1; }
  return 0;
}

void fvla_destroy(fvla_engine* e) { delete e; }

int fvla_load_tensor(fvla_engine* e, const char* name, const void* data, int32_t dtype, int32_t ndim,
                     const int64_t* shape) {
  if (e == nullptr) { set_error("null engine"); return 2; }
  return e->impl.load_tensor(name, data, dtype, ndim, shape);
}

int fvla_missing_tensors(fvla_engine* e, char* buf, int64_t buf_len, int32_t* n_missing) {
  if (e == nullptr) { set_error("null engine"); return 2; }
  std::vector<std::string> miss;
  e->impl.missing(&miss);
  if (n_missing) *n_missing = static_cast<int32_t>(miss.size());
  if (buf != nullptr && buf_len > 0) {
    std::string joined;
    for (auto& m : miss) { joined += m; joined += '\n'; }
    const size_t n = std::min<size_t>(joined.size(), static_cast<size_t>(buf_len - 1));
    std::memcpy(buf, joined.data(), n);
    buf[n] = '\0';
  }
  return 0;
}

int fvla_finalize(fvla_engine* e) {
  if (e == nullptr) { set_error("null engine"); return 2; }
  return e->impl.finalize();
}
int fvla_reserve(fvla_engine* e, int32_t batch, int32_t n_tokens) {
  if (e == nullptr) { set_error("null engine"); return 2; }
  return e->impl.reserve(batch, n_tokens);
}
int64_t fvla_workspace_bytes(fvla_engine* e) { return e ? static_cast<int64_t>(e->impl.workspace_bytes()) : 0; }
int64_t fvla_weight_bytes(fvla_engine* e) { return e ? static_cast<int64_t>(e->impl.weight_bytes) : 0; }

int fvla_forward(fvla_engine* e, const fvla_forward_args* args, void* stream) {
  if (e == nullptr || args == nullptr) { set_error("null argument"); return 2; }
  return e->impl.forward(*args, static_cast<cudaStream_t>(stream));
}
int64_t fvla_last_launch_count(fvla_engine* e) { return e ? e->impl.launches : 0; }
double fvla_last_forward_flops(fvla_engine* e) { return e ? e->impl.flops : 0.0; }
int fvla_set_tap(fvla_engine* e, int32_t stage, void* dst, int64_t cap) {
  if (e == nullptr) { set_error("null engine"); return 2; }
  return e->impl.set_tap(stage, dst, cap);
}
int fvla_merged_len(fvla_engine* e) { return e ? e->impl.merged_len : 0; }
int fvla_set_io_normalization(fvla_engine* e, const float* state_mean, const float* state_inv_std,
                              const float* action_scale, const float* action_shift) {
  if (e == nullptr) { set_error("null engine"); return 2; }
  return e->impl.set_io_normalization(state_mean, state_inv_std, action_scale, action_shift);
}
int fvla_set_profile(fvla_engine* e, int32_t on) {
  if (e == nullptr) { set_error("null engine"); return 2; }
  e->impl.set_profile(on != 0);
  return 0;
}
int fvla_profile_report(fvla_engine* e, char* buf, int64_t buf_len) {
  if (e == nullptr || buf == nullptr || buf_len <= 0) { set_error("null argument"); return 2; }
  std::string csv;
  if (int rc = e->impl.profile_report(&csv)) return rc;
  const size_t n = std::min<size_t>(csv.size(), static_cast<size_t>(buf_len - 1));
  std::memcpy(buf, csv.data(), n);
  buf[n] = '\0';
  return 0;
}

// ---- single-kernel entry points ----
int fvla_op_gemm(int32_t dtype, const void* A, int32_t lda, const void* W, int32_t ldw, void* D,
                 int32_t ldd, int32_t M, int32_t N, int32_t K, const float* bias,
                 const float* row_scale, const void* resid, int32_t ldr, int32_t act, int32_t swiglu,
                 int32_t block_n, void* stream) {
  fvla::GemmArgs g;
  g.A = A; g.lda = lda; g.W = W; g.ldw = ldw; g.D = D; g.ldd = ldd;
  g.M = M; g.N = N; g.K = K; g.bias = bias; g.row_scale = row_scale;
  g.resid = resid; g.ldr = ldr; g.act = act; g.swiglu = swiglu & 1; g.ab_f16 = (swiglu >> 1) & 1;
  g.out_f32 = (swiglu >> 2) & 1;
  g.split_k = (swiglu >> 8) & 0xff;  // bits 8..15 of the flag word: split-K factor of the reduce-add epilogue (0 = none)
  g.block_n = block_n;
  return fvla::gemm(dtype, g, static_cast<cudaStream_t>(stream));
}

int fvla_op_preprocess(int32_t dtype, const void* src, int32_t src_dtype, int32_t nhwc, int32_t B,
                       int32_t C, int32_t h, int32_t w, int32_t S, int32_t letterbox, float pad_value,
                       float scale, int32_t normalize, const float* mean3, const float* inv_std3,
                       void* dst, void* stream) {
  fvla::PreprocessArgs a;
  a.src = src; a.src_dtype = src_dtype; a.src_nhwc = nhwc; a.B = B; a.C = C; a.h = h; a.w = w;
  a.S = S; a.letterbox = letterbox; a.pad_value = pad_value; a.scale = scale; a.normalize = normalize;
  for (int i = 0; i < 3; ++i) {
    a.mean[i] = mean3 ? mean3[i] : 0.f;
    a.inv_std[i] = inv_std3 ? inv_std3[i] : 1.f;
  }
  a.dst = dst;
  return fvla::preprocess_images(dtype, a, static_cast<cudaStream_t>(stream));
}

int fvla_op_stem_conv(int32_t dtype, const void* in, const float* w_packed, const float* bias,
                      void* out, int32_t B, int32_t H, int32_t W, int32_t Cout, void* stream) {
  return fvla::stem_conv3x3_s2(dtype, in, w_packed, bias, out, B, H, W, Cout,
                               static_cast<cudaStream_t>(stream));
}

int fvla_op_dwconv(int32_t dtype, const void* in, const float* w_packed, const float* bias, void* out,
                   int32_t B, int32_t H, int32_t W, int32_t Cin, int32_t mult, int32_t ksize,
                   int32_t stride, int32_t act, void* stream) {
  return fvla::dwconv(dtype, in, w_packed, bias, out, B, H, W, Cin, mult, ksize, stride, act,
                      static_cast<cudaStream_t>(stream));
}

int fvla_op_se_gelu(int32_t dtype, const void* x, void* out, int32_t B, int32_t HW, int32_t C,
                    int32_t Cr, const float* w1, const float* b1, const float* w2, const float* b2,
                    float* scratch_mean, float* scratch_gate, void* stream) {
  return fvla::se_gelu(dtype, x, out, B, HW, C, Cr, w1, b1, w2, b2, scratch_mean, scratch_gate,
                       static_cast<cudaStream_t>(stream));
}

int fvla_op_attention(int32_t dtype, int32_t impl, const void* q, const void* k, const void* v,
                      int32_t ld_qkv, void* o, int32_t ld_o, int32_t B, int32_t N, int32_t heads_q,
                      int32_t heads_kv, int32_t head_dim, float scale, int32_t causal,
                      const float* rope_cos, const float* rope_sin, void* stream) {
  fvla::AttnArgs a;
  a.q = q; a.k = k; a.v = v; a.ld_qkv = ld_qkv; a.o = o; a.ld_o = ld_o; a.B = B; a.N = N;
  a.heads_q = heads_q; a.heads_kv = heads_kv; a.head_dim = head_dim; a.scale = scale;
  a.causal = causal; a.rope_cos = rope_cos; a.rope_sin = rope_sin;
  if (impl == 1) return fvla::attention_simt(dtype, a, static_cast<cudaStream_t>(stream));
  return fvla::attention(dtype, a, static_cast<cudaStream_t>(stream));
}

int fvla_op_rmsnorm(int32_t dtype, const void* x, const float* weight, void* out, int32_t rows,
                    int32_t H, float eps, void* stream) {
  return fvla::rmsnorm(dtype, x, weight, out, rows, H, eps, static_cast<cudaStream_t>(stream));
}

int fvla_op_layernorm_rows(int32_t dtype, const void* x, void* out, int32_t rows, int32_t C, float eps, void* stream) {
  return fvla::layernorm_rows(dtype, x, out, rows, C, eps, static_cast<cudaStream_t>(stream));
}

int fvla_op_ffn_fused(const void* x, const void* w1_half, const float* b1_half, const void* w2,
                      const float* b2, const void* resid, void* out, int32_t M, int32_t C, int32_t hidden,
                      void* stream) {
  fvla::FfnFusedArgs a;
  a.x = x; a.w1 = w1_half; a.b1 = b1_half; a.w2 = w2; a.b2 = b2; a.resid = resid; a.out = out;
  a.M = M; a.C = C; a.hidden = hidden;
  return fvla::ffn_fused(a, static_cast<cudaStream_t>(stream));
}

int64_t fvla_head_train_scratch_floats(int32_t B, int32_t H, int32_t S, int32_t Hd, int32_t F, int32_t A) {
  return static_cast<int64_t>(fvla::head_train_scratch_floats(B, H, S, Hd, F, A));
}

int fvla_head_forward_backward(int32_t B, int32_t H, int32_t S, int32_t Hd, int32_t F, int32_t A,
                               const float* const* params, const float* pooled, const float* states,
                               const float* target, const uint8_t* keep_mask, float drop_p, float* grads, float* loss,
                               float* actions, float* scratch, int64_t scratch_floats, void* stream) {
  if (params == nullptr || pooled == nullptr || states == nullptr || target == nullptr || grads == nullptr ||
      loss == nullptr) {
    set_error("fvla_head_forward_backward: null argument");
    return 2;
  }
  for (int i = 0; i < 12; ++i)
    if (params[i] == nullptr) { set_error("fvla_head_forward_backward: null parameter tensor"); return 2; }
  fvla::HeadTrainArgs t;
  t.B = B; t.H = H; t.S = S; t.Hd = Hd; t.F = F; t.A = A;
  t.w.ln_s_w = params[0]; t.w.ln_s_b = params[1]; t.w.w_state = params[2]; t.w.b_state = params[3];
  t.w.w_f0 = params[4]; t.w.b_f0 = params[5]; t.w.ln_f_w = params[6]; t.w.ln_f_b = params[7];
  t.w.w_f4 = params[8]; t.w.b_f4 = params[9]; t.w.w_act = params[10]; t.w.b_act = params[11];
  t.w.H = H; t.w.S = S; t.w.Hd = Hd; t.w.F = F; t.w.A = A;
  t.pooled = pooled; t.states = states; t.target = target; t.keep_mask = keep_mask; t.drop_p = drop_p;
  t.grads = grads; t.loss = loss; t.actions = actions; t.scratch = scratch;
  t.scratch_floats = scratch_floats > 0 ? static_cast<size_t>(scratch_floats) : 0;
  return fvla::head_train_step(t, static_cast<cudaStream_t>(stream));
}

int fvla_op_convert(int32_t src_dtype, const void* src, int32_t dst_dtype, void* dst, int64_t n,
                    void* stream) {
  return fvla::convert(src_dtype, src, dst_dtype, dst, n, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
