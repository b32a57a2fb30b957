// Engine: weight import (HF state-dict names), BatchNorm / layer-scale folding, repacking, workspace
// and the launch sequence of the FastVLA forward.  See engine.h / include/fvla.h.
#include "engine.h"

#include <cuda_fp16.h>

#include <cstdlib>
#include "common.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>

namespace fvla {

namespace {
const char* kVis = "backbone.model.model.vision_tower.vision_tower.model.";
const char* kProj = "backbone.model.model.mm_projector.";
const char* kLlm = "backbone.model.model.";
constexpr float kBnEps = 1e-5f;  // torch.nn.BatchNorm2d default
constexpr float kLnChannelEps = 1e-5f;  // LayerNormChannel default [EXT mci.py]

std::string S(const char* p, const std::string& rest) { return std::string(p) + rest; }
}  // namespace

Engine::Engine(const fvla_config& c) : cfg(c) {}

Engine::~Engine() {
  clear_graphs();
  if (graph_stream_ != nullptr) cudaStreamDestroy(graph_stream_);
  for (void* p : dev_allocs_) cudaFree(p);
  for (auto& kv : ws_.bufs) cudaFree(kv.second.first);
  for (auto& p : prof_) { cudaEventDestroy(p.e0); cudaEventDestroy(p.e1); }
  for (auto e : ev_pool_) cudaEventDestroy(e);
  for (auto& hs : host_ring_) {
    if (hs.ev != nullptr) cudaEventDestroy(hs.ev);
    if (hs.p != nullptr) cudaFreeHost(hs.p);
  }
}

// ---------------------------------------------------------------------------------------------
// Per-launch profiling
// ---------------------------------------------------------------------------------------------
cudaEvent_t Engine::get_event() {
  if (!ev_pool_.empty()) { cudaEvent_t e = ev_pool_.back(); ev_pool_.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void Engine::prof_begin(cudaStream_t s) {
  if (!profile_) return;
  prof_e0_ = get_event();
  cudaEventRecord(prof_e0_, s);
}
void Engine::prof_end(const std::string& label, double fl, double by, cudaStream_t s) {
  if (!profile_) return;
  cudaEvent_t e1 = get_event();
  cudaEventRecord(e1, s);
  prof_.push_back({label, prof_e0_, e1, fl, by});
  prof_e0_ = nullptr;
}
int Engine::profile_report(std::string* csv) {
  FVLA_CUDA_CHECK(cudaDeviceSynchronize());
  struct Agg { long long n = 0; double ms = 0, fl = 0, by = 0; };
  std::map<std::string, Agg> agg;
  std::vector<std::string> order;
  for (auto& p : prof_) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, p.e0, p.e1);
    if (agg.find(p.label) == agg.end()) order.push_back(p.label);
    Agg& a = agg[p.label];
    a.n += 1; a.ms += ms; a.fl += p.flops; a.by += p.bytes;
    ev_pool_.push_back(p.e0);
    ev_pool_.push_back(p.e1);
  }
  prof_.clear();
  csv->clear();
  *csv += "label,count,total_ms,flops,bytes\n";
  char line[512];
  for (auto& k : order) {
    const Agg& a = agg[k];
    snprintf(line, sizeof(line), "%s,%lld,%.6f,%.6e,%.6e\n", k.c_str(), a.n, a.ms, a.fl, a.by);
    *csv += line;
  }
  return 0;
}

int Engine::n_img_tokens() const {
  int side = cfg.image_size / 4;  // stem: two stride-2 convs
  for (int i = 0; i + 1 < cfg.vis_num_stages; ++i) side /= 2;
  return side * side;
}
int Engine::mm_hidden() const { return cfg.vis_dims[cfg.vis_num_stages - 1] * 2; }  // cls_ratio 2.0

// ---------------------------------------------------------------------------------------------
// Tensor staging
// ---------------------------------------------------------------------------------------------
int Engine::load_tensor(const char* name, const void* data, int dtype, int ndim,
                        const int64_t* shape) {
  FVLA_REQUIRE(name != nullptr && data != nullptr && ndim >= 0 && ndim <= 8, "bad tensor");
  HostTensor t;
  t.shape.assign(shape, shape + ndim);
  const int64_t n = t.numel();
  t.data.resize(static_cast<size_t>(n));
  if (dtype == FVLA_F32) {
    std::memcpy(t.data.data(), data, static_cast<size_t>(n) * 4);
  } else if (dtype == FVLA_BF16) {
    const uint16_t* s = static_cast<const uint16_t*>(data);
    for (int64_t i = 0; i < n; ++i) {
      const uint32_t u = static_cast<uint32_t>(s[i]) << 16;
      std::memcpy(&t.data[static_cast<size_t>(i)], &u, 4);
    }
  } else {
    set_error("load_tensor: dtype must be fp32 or bf16");
    return 2;
  }
  if (finalized_) return update_head_tensor(name, t);
  host_[name] = std::move(t);
  return 0;
}

const HostTensor* Engine::find(const std::string& name) const {
  auto it = host_.find(name);
  return it == host_.end() ? nullptr : &it->second;
}

int Engine::need(const std::string& name, const HostTensor** out,
                 std::initializer_list<int64_t> shape) {
  const HostTensor* t = find(name);
  if (t == nullptr) { set_error("missing tensor: " + name); return 3; }
  int64_t want = 1;
  for (auto d : shape) want *= d;
  if (t->numel() != want) {
    set_error("tensor " + name + " has " + std::to_string(t->numel()) + " elements, expected " +
              std::to_string(want));
    return 3;
  }
  *out = t;
  return 0;
}

std::vector<std::string> Engine::required_names() const {
  std::vector<std::string> r;
  auto wb = [&](const std::string& p) { r.push_back(p + ".weight"); r.push_back(p + ".bias"); };
  auto bn = [&](const std::string& p) {
    wb(p); r.push_back(p + ".running_mean"); r.push_back(p + ".running_var");
  };
  for (int i = 0; i < 3; ++i) wb(S(kVis, "patch_embed." + std::to_string(i) + ".reparam_conv"));
  int idx = 0;
  for (int i = 0; i < cfg.vis_num_stages; ++i) {
    if (cfg.vis_pos_emb[i]) wb(S(kVis, "network." + std::to_string(idx++) + ".reparam_conv"));
    const int st = idx++;
    for (int j = 0; j < cfg.vis_layers[i]; ++j) {
      const std::string b = S(kVis, "network." + std::to_string(st) + "." + std::to_string(j));
      if (cfg.vis_attention[i]) {
        // pre-attention norm: BatchNorm2d (running stats present) or LayerNormChannel (weight/bias only) — the
        // checkpoint's key set decides (pack_vision)
        if (host_.find(b + ".norm.running_mean") != host_.end() || host_.find(b + ".norm.running_var") != host_.end())
          bn(b + ".norm");
        else
          wb(b + ".norm");
        r.push_back(b + ".token_mixer.qkv.weight");
        wb(b + ".token_mixer.proj");
        r.push_back(b + ".layer_scale_1");
        r.push_back(b + ".layer_scale_2");
      } else {
        wb(b + ".token_mixer.reparam_conv");
        r.push_back(b + ".layer_scale");
      }
      r.push_back(b + ".convffn.conv.conv.weight");
      bn(b + ".convffn.conv.bn");
      wb(b + ".convffn.fc1");
      wb(b + ".convffn.fc2");
    }
    if (i + 1 < cfg.vis_num_stages) {
      const std::string p = S(kVis, "network." + std::to_string(idx++));
      wb(p + ".proj.0.lkb_reparam");
      wb(p + ".proj.1.reparam_conv");
    }
  }
  wb(S(kVis, "conv_exp.reparam_conv"));
  wb(S(kVis, "conv_exp.se.reduce"));
  wb(S(kVis, "conv_exp.se.expand"));
  wb(S(kProj, "0"));
  wb(S(kProj, "2"));
  r.push_back(S(kLlm, "embed_tokens.weight"));
  for (int l = 0; l < cfg.n_layers; ++l) {
    const std::string b = S(kLlm, "layers." + std::to_string(l));
    r.push_back(b + ".input_layernorm.weight");
    wb(b + ".self_attn.q_proj"); wb(b + ".self_attn.k_proj"); wb(b + ".self_attn.v_proj");
    r.push_back(b + ".self_attn.o_proj.weight");
    r.push_back(b + ".post_attention_layernorm.weight");
    r.push_back(b + ".mlp.gate_proj.weight");
    r.push_back(b + ".mlp.up_proj.weight");
    r.push_back(b + ".mlp.down_proj.weight");
  }
  r.push_back(S(kLlm, "norm.weight"));
  if (cfg.state_dim > 0) {  // state_dim 0 = backbone-only engine (FastVLMBackbone used on its own)
    wb("state_projection.0"); wb("state_projection.1");
    wb("fusion.0"); wb("fusion.1"); wb("fusion.4");
    wb("action_head");
  }
  return r;
}

int Engine::missing(std::vector<std::string>* out) {
  out->clear();
  for (const auto& n : required_names())
    if (host_.find(n) == host_.end()) out->push_back(n);
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Uploads
// ---------------------------------------------------------------------------------------------
float* Engine::upload_f32(const std::vector<float>& v) {
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(v.size() * 4, 16);
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  cudaMemcpy(p, v.data(), v.size() * 4, cudaMemcpyHostToDevice);
  dev_allocs_.push_back(p);
  weight_bytes += bytes;
  return static_cast<float*>(p);
}

void* Engine::upload_act(const std::vector<float>& v) {
  if (cfg.dtype == FVLA_F32) return upload_f32(v);
  std::vector<__nv_bfloat16> h(v.size());
  for (size_t i = 0; i < v.size(); ++i) h[i] = __float2bfloat16_rn(v[i]);
  void* p = nullptr;
  const size_t bytes = std::max<size_t>(h.size() * 2, 16);
  if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  cudaMemcpy(p, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  dev_allocs_.push_back(p);
  weight_bytes += bytes;
  return p;
}

int Engine::make_gemm(GemmW* g, const std::vector<float>& w_in, int N, int K,
                      const std::vector<float>* bias_in, bool gelu_half) {
  FVLA_REQUIRE(static_cast<int64_t>(w_in.size()) == static_cast<int64_t>(N) * K, "gemm weight size");
  // GELU GEMMs produce x/2 (scaling by a power of two is exact) so the epilogue saves two multiplies
  std::vector<float> w_scaled, b_scaled;
  const std::vector<float>* bias = bias_in;
  if (gelu_half) {
    w_scaled.resize(w_in.size());
    for (size_t i = 0; i < w_in.size(); ++i) w_scaled[i] = 0.5f * w_in[i];
    if (bias_in != nullptr) {
      b_scaled.resize(bias_in->size());
      for (size_t i = 0; i < bias_in->size(); ++i) b_scaled[i] = 0.5f * (*bias_in)[i];
      bias = &b_scaled;
    }
  }
  const std::vector<float>& w = gelu_half ? w_scaled : w_in;
  g->half_in = gelu_half;
  g->N = N; g->K = K;
  g->w = upload_act(w);
  FVLA_REQUIRE(g->w != nullptr, "cudaMalloc failed for a GEMM weight");
  g->bias = nullptr;
  if (bias != nullptr) {
    FVLA_REQUIRE(static_cast<int>(bias->size()) == N, "gemm bias size");
    g->bias = upload_f32(*bias);
    FVLA_REQUIRE(g->bias != nullptr, "cudaMalloc failed for a bias");
  }
  return 0;
}

int Engine::make_dw(DwW* d, const std::vector<float>& w, const std::vector<float>& bias, int cin,
                    int mult, int k, int stride, int act) {
  const int cout = cin * mult;
  FVLA_REQUIRE(static_cast<int>(w.size()) == cout * k * k, "depthwise weight size");
  FVLA_REQUIRE(static_cast<int>(bias.size()) == cout, "depthwise bias size");
  std::vector<float> packed(static_cast<size_t>(k) * k * cout);
  for (int o = 0; o < cout; ++o)
    for (int t = 0; t < k * k; ++t) packed[static_cast<size_t>(t) * cout + o] = w[static_cast<size_t>(o) * k * k + t];
  d->w = upload_f32(packed);
  d->bias = upload_f32(bias);
  FVLA_REQUIRE(d->w != nullptr && d->bias != nullptr, "cudaMalloc failed for a depthwise conv");
  d->cin = cin; d->mult = mult; d->k = k; d->stride = stride; d->act = act;
  if (cfg.dtype == FVLA_BF16 && k == 7 && stride == 1 && mult == 1 && act == ACT_NONE && cin % 32 == 0) {
    void* p = nullptr;
    FVLA_CUDA_CHECK(cudaMalloc(&p, dwconv7_wtab_bytes(cin)));
    dev_allocs_.push_back(p);
    weight_bytes += dwconv7_wtab_bytes(cin);
    d->wtab = static_cast<uint32_t*>(p);
    if (int rc = dwconv7_mma_prepare(d->w, cin, d->wtab, nullptr)) return rc;
  } else if (cfg.dtype == FVLA_BF16 && k == 7 && stride == 2 && mult == 2 && cin % 16 == 0) {
    void* p = nullptr;   // the stride-2 tensor-core path's table (dwconv() falls back by shape, the table is harmless then)
    FVLA_CUDA_CHECK(cudaMalloc(&p, dwconv7_wtab_bytes(cin)));
    dev_allocs_.push_back(p);
    weight_bytes += dwconv7_wtab_bytes(cin);
    d->wtab = static_cast<uint32_t*>(p);
    if (int rc = dwconv7_s2m2_mma_prepare(d->w, cin, d->wtab, nullptr)) return rc;
  }
  return 0;
}

// BatchNorm2d (eval) as per-channel affine: y = s*x + t
static void bn_affine(const HostTensor& w, const HostTensor& b, const HostTensor& mean,
                      const HostTensor& var, std::vector<double>* s, std::vector<double>* t) {
  const size_t c = w.data.size();
  s->resize(c); t->resize(c);
  for (size_t i = 0; i < c; ++i) {
    const double sc = static_cast<double>(w.data[i]) / std::sqrt(static_cast<double>(var.data[i]) + kBnEps);
    (*s)[i] = sc;
    (*t)[i] = static_cast<double>(b.data[i]) - static_cast<double>(mean.data[i]) * sc;
  }
}

int Engine::pack_vision() {
  const HostTensor *w, *b;
  const int d0 = cfg.vis_dims[0];
  // ---- stem ----
  if (int rc = need(S(kVis, "patch_embed.0.reparam_conv.weight"), &w, {d0, 3, 3, 3})) return rc;
  if (int rc = need(S(kVis, "patch_embed.0.reparam_conv.bias"), &b, {d0})) return rc;
  {
    std::vector<float> packed(27 * static_cast<size_t>(d0));
    for (int o = 0; o < d0; ++o)
      for (int ci = 0; ci < 3; ++ci)
        for (int ky = 0; ky < 3; ++ky)
          for (int kx = 0; kx < 3; ++kx)
            packed[static_cast<size_t>((ky * 3 + kx) * 3 + ci) * d0 + o] =
                w->data[((static_cast<size_t>(o) * 3 + ci) * 3 + ky) * 3 + kx];
    stem0_w_ = upload_f32(packed);
    stem0_b_ = upload_f32(b->data);
    FVLA_REQUIRE(stem0_w_ && stem0_b_, "cudaMalloc failed (stem)");
    if (cfg.dtype == FVLA_BF16) {
      // im2col form for the tensor cores: [d0][32] with column (ky*3+kx)*3+ci, columns 27..31 zero
      std::vector<float> wg(32 * static_cast<size_t>(d0), 0.f);
      for (int o = 0; o < d0; ++o)
        for (int t = 0; t < 27; ++t) wg[static_cast<size_t>(o) * 32 + t] = packed[static_cast<size_t>(t) * d0 + o];
      if (int rc = make_gemm(&stem0_gemm_, wg, d0, 32, &b->data, true)) return rc;
      if (stem_fused_supported(cfg.dtype, cfg.image_size, d0)) {
        std::vector<uint32_t> tab(stem_fused_btab_words(d0));
        stem_fused_build_btab(packed.data(), d0, tab.data());
        void* p = nullptr;
        FVLA_CUDA_CHECK(cudaMalloc(&p, tab.size() * 4));
        FVLA_CUDA_CHECK(cudaMemcpy(p, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
        dev_allocs_.push_back(p);
        stem_btab_ = static_cast<uint32_t*>(p);
        std::vector<float> bh(b->data);
        for (auto& v : bh) v *= 0.5f;
        stem0_bh_ = upload_f32(bh);
        FVLA_REQUIRE(stem0_bh_ != nullptr, "cudaMalloc failed (stem bias)");
      }
    }
  }
  if (int rc = need(S(kVis, "patch_embed.1.reparam_conv.weight"), &w, {d0, 1, 3, 3})) return rc;
  if (int rc = need(S(kVis, "patch_embed.1.reparam_conv.bias"), &b, {d0})) return rc;
  if (int rc = make_dw(&stem1_, w->data, b->data, d0, 1, 3, 2, ACT_GELU)) return rc;
  if (int rc = need(S(kVis, "patch_embed.2.reparam_conv.weight"), &w, {d0, d0, 1, 1})) return rc;
  if (int rc = need(S(kVis, "patch_embed.2.reparam_conv.bias"), &b, {d0})) return rc;
  if (int rc = make_gemm(&stem2_, w->data, d0, d0, &b->data, true)) return rc;

  // ---- stages ----
  auto pack_ffn = [&](const std::string& base, int d, const HostTensor* ls, VisBlock* blk) -> int {
    const HostTensor *cw, *bw, *bb, *bm, *bv, *f1w, *f1b, *f2w, *f2b;
    const int hd = d * cfg.vis_mlp_ratio;
    if (int rc = need(base + ".convffn.conv.conv.weight", &cw, {d, 1, 7, 7})) return rc;
    if (int rc = need(base + ".convffn.conv.bn.weight", &bw, {d})) return rc;
    if (int rc = need(base + ".convffn.conv.bn.bias", &bb, {d})) return rc;
    if (int rc = need(base + ".convffn.conv.bn.running_mean", &bm, {d})) return rc;
    if (int rc = need(base + ".convffn.conv.bn.running_var", &bv, {d})) return rc;
    if (int rc = need(base + ".convffn.fc1.weight", &f1w, {hd, d, 1, 1})) return rc;
    if (int rc = need(base + ".convffn.fc1.bias", &f1b, {hd})) return rc;
    if (int rc = need(base + ".convffn.fc2.weight", &f2w, {d, hd, 1, 1})) return rc;
    if (int rc = need(base + ".convffn.fc2.bias", &f2b, {d})) return rc;
    std::vector<double> s, t;
    bn_affine(*bw, *bb, *bm, *bv, &s, &t);
    std::vector<float> dw(cw->data.size()), dwb(static_cast<size_t>(d));
    for (int c = 0; c < d; ++c) {
      for (int i = 0; i < 49; ++i)
        dw[static_cast<size_t>(c) * 49 + i] = static_cast<float>(cw->data[static_cast<size_t>(c) * 49 + i] * s[c]);
      dwb[c] = static_cast<float>(t[c]);
    }
    if (int rc = make_dw(&blk->ffn_dw, dw, dwb, d, 1, 7, 1, ACT_NONE)) return rc;
    if (int rc = make_gemm(&blk->fc1, f1w->data, hd, d, &f1b->data, true)) return rc;
    // x + ls * (W h + b)  ==  x + (ls.W) h + ls.b
    std::vector<float> w2(f2w->data.size()), b2(static_cast<size_t>(d));
    for (int n = 0; n < d; ++n) {
      const float l = ls->data[n];
      for (int k = 0; k < hd; ++k) w2[static_cast<size_t>(n) * hd + k] = f2w->data[static_cast<size_t>(n) * hd + k] * l;
      b2[n] = f2b->data[n] * l;
    }
    if (int rc = make_gemm(&blk->fc2, w2, d, hd, &b2)) return rc;
    if (cfg.dtype == FVLA_BF16) {
      // the ConvFFN hidden tensor is fp16 (on chip in the fused kernel, in HBM otherwise): GELU runs on packed
      // half pairs and fc2 is an fp16 x fp16 MMA
      std::vector<__half> h(w2.size());
      for (size_t i = 0; i < w2.size(); ++i) h[i] = __float2half_rn(w2[i]);
      void* p = nullptr;
      FVLA_CUDA_CHECK(cudaMalloc(&p, h.size() * 2));
      FVLA_CUDA_CHECK(cudaMemcpy(p, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
      dev_allocs_.push_back(p);
      weight_bytes += h.size() * 2;
      blk->fc2.w_f16 = p;
    }
    return 0;
  };

  stages_.clear();
  stages_.resize(cfg.vis_num_stages);
  int idx = 0;
  for (int i = 0; i < cfg.vis_num_stages; ++i) {
    VisStage& st = stages_[i];
    const int d = cfg.vis_dims[i];
    st.dim = d;
    if (cfg.vis_pos_emb[i]) {
      const std::string p = S(kVis, "network." + std::to_string(idx++) + ".reparam_conv");
      if (int rc = need(p + ".weight", &w, {d, 1, 7, 7})) return rc;
      if (int rc = need(p + ".bias", &b, {d})) return rc;
      st.has_cpe = true;
      if (int rc = make_dw(&st.cpe, w->data, b->data, d, 1, 7, 1, ACT_NONE)) return rc;
    }
    const int sidx = idx++;
    st.blocks.resize(cfg.vis_layers[i]);
    for (int j = 0; j < cfg.vis_layers[i]; ++j) {
      VisBlock& blk = st.blocks[j];
      const std::string base = S(kVis, "network." + std::to_string(sidx) + "." + std::to_string(j));
      if (cfg.vis_attention[i]) {
        blk.attn = true;
        FVLA_REQUIRE(d % cfg.vis_head_dim == 0, "attention stage width must be a multiple of head_dim");
        const HostTensor *nw, *nb, *nm = nullptr, *nv = nullptr, *qw, *pw, *pb, *ls1, *ls2;
        if (int rc = need(base + ".norm.weight", &nw, {d})) return rc;
        if (int rc = need(base + ".norm.bias", &nb, {d})) return rc;
        // BatchNorm2d carries running statistics and folds into qkv completely; LayerNormChannel [EXT mci.py]
        // (per-token statistics over the channels) keeps a run-time normalisation kernel and folds only its affine
        blk.attn_layernorm = find(base + ".norm.running_mean") == nullptr;
        if (!blk.attn_layernorm) {
          if (int rc = need(base + ".norm.running_mean", &nm, {d})) return rc;
          if (int rc = need(base + ".norm.running_var", &nv, {d})) return rc;
        }
        if (int rc = need(base + ".token_mixer.qkv.weight", &qw, {3 * d, d})) return rc;
        if (int rc = need(base + ".token_mixer.proj.weight", &pw, {d, d})) return rc;
        if (int rc = need(base + ".token_mixer.proj.bias", &pb, {d})) return rc;
        if (int rc = need(base + ".layer_scale_1", &ls1, {d})) return rc;
        if (int rc = need(base + ".layer_scale_2", &ls2, {d})) return rc;
        std::vector<double> s, t;
        if (blk.attn_layernorm) {
          s.assign(nw->data.begin(), nw->data.end());
          t.assign(nb->data.begin(), nb->data.end());
        } else {
          bn_affine(*nw, *nb, *nm, *nv, &s, &t);
        }
        // qkv(norm(x)) = (W diag(s)) n + W t   (n = x for BatchNorm, the normalised row for LayerNormChannel)
        std::vector<float> qf(qw->data.size()), qb(static_cast<size_t>(3) * d);
        for (int n = 0; n < 3 * d; ++n) {
          double acc = 0.0;
          for (int k = 0; k < d; ++k) {
            const double wv = qw->data[static_cast<size_t>(n) * d + k];
            qf[static_cast<size_t>(n) * d + k] = static_cast<float>(wv * s[k]);
            acc += wv * t[k];
          }
          qb[n] = static_cast<float>(acc);
        }
        if (int rc = make_gemm(&blk.qkv, qf, 3 * d, d, &qb)) return rc;
        std::vector<float> pf(pw->data.size()), pbf(static_cast<size_t>(d));
        for (int n = 0; n < d; ++n) {
          const float l = ls1->data[n];
          for (int k = 0; k < d; ++k) pf[static_cast<size_t>(n) * d + k] = pw->data[static_cast<size_t>(n) * d + k] * l;
          pbf[n] = pb->data[n] * l;
        }
        if (int rc = make_gemm(&blk.proj, pf, d, d, &pbf)) return rc;
        if (int rc = pack_ffn(base, d, ls2, &blk)) return rc;
      } else {
        const HostTensor *mw, *mb, *ls;
        if (int rc = need(base + ".token_mixer.reparam_conv.weight", &mw, {d, 1, 3, 3})) return rc;
        if (int rc = need(base + ".token_mixer.reparam_conv.bias", &mb, {d})) return rc;
        if (int rc = need(base + ".layer_scale", &ls, {d})) return rc;
        if (int rc = make_dw(&blk.mixer, mw->data, mb->data, d, 1, 3, 1, ACT_NONE)) return rc;
        if (int rc = pack_ffn(base, d, ls, &blk)) return rc;
      }
    }
    if (i + 1 < cfg.vis_num_stages) {
      const int d2 = cfg.vis_dims[i + 1];
      FVLA_REQUIRE(d2 == 2 * d, "patch-embed grouped conv expects the next stage to double the width");
      const std::string p = S(kVis, "network." + std::to_string(idx++));
      const HostTensor *lw, *lb, *pw, *pb;
      if (int rc = need(p + ".proj.0.lkb_reparam.weight", &lw, {d2, 1, 7, 7})) return rc;
      if (int rc = need(p + ".proj.0.lkb_reparam.bias", &lb, {d2})) return rc;
      if (int rc = need(p + ".proj.1.reparam_conv.weight", &pw, {d2, d2, 1, 1})) return rc;
      if (int rc = need(p + ".proj.1.reparam_conv.bias", &pb, {d2})) return rc;
      st.has_down = true;
      if (int rc = make_dw(&st.down_dw, lw->data, lb->data, d, 2, 7, 2, ACT_GELU)) return rc;
      if (int rc = make_gemm(&st.down_pw, pw->data, d2, d2, &pb->data, true)) return rc;
    }
  }
  // ---- conv_exp ----
  const int dl = cfg.vis_dims[cfg.vis_num_stages - 1], ce = 2 * dl, cr = cfg.vis_se_reduced;
  if (int rc = need(S(kVis, "conv_exp.reparam_conv.weight"), &w, {ce, 1, 3, 3})) return rc;
  if (int rc = need(S(kVis, "conv_exp.reparam_conv.bias"), &b, {ce})) return rc;
  if (int rc = make_dw(&exp_dw_, w->data, b->data, dl, 2, 3, 1, ACT_NONE)) return rc;
  if (int rc = need(S(kVis, "conv_exp.se.reduce.weight"), &w, {cr, ce, 1, 1})) return rc;
  if (int rc = need(S(kVis, "conv_exp.se.reduce.bias"), &b, {cr})) return rc;
  se_w1_ = upload_f32(w->data); se_b1_ = upload_f32(b->data);
  if (int rc = need(S(kVis, "conv_exp.se.expand.weight"), &w, {ce, cr, 1, 1})) return rc;
  if (int rc = need(S(kVis, "conv_exp.se.expand.bias"), &b, {ce})) return rc;
  se_w2_ = upload_f32(w->data); se_b2_ = upload_f32(b->data);
  FVLA_REQUIRE(se_w1_ && se_b1_ && se_w2_ && se_b2_, "cudaMalloc failed (SE)");
  // ---- projector ----
  const int H = cfg.hidden;
  if (int rc = need(S(kProj, "0.weight"), &w, {H, ce})) return rc;
  if (int rc = need(S(kProj, "0.bias"), &b, {H})) return rc;
  if (int rc = make_gemm(&proj0_, w->data, H, ce, &b->data, true)) return rc;
  if (int rc = need(S(kProj, "2.weight"), &w, {H, H})) return rc;
  if (int rc = need(S(kProj, "2.bias"), &b, {H})) return rc;
  return make_gemm(&proj2_, w->data, H, H, &b->data);
}

int Engine::pack_decoder() {
  const int H = cfg.hidden, nq = cfg.n_q_heads, nkv = cfg.n_kv_heads, hd = cfg.head_dim;
  const int I = cfg.intermediate;
  const HostTensor* t;
  if (int rc = need(S(kLlm, "embed_tokens.weight"), &t, {cfg.vocab, H})) return rc;
  embed_ = upload_act(t->data);
  FVLA_REQUIRE(embed_ != nullptr, "cudaMalloc failed (embedding table)");
  layers_.clear();
  layers_.resize(cfg.n_layers);
  for (int l = 0; l < cfg.n_layers; ++l) {
    DecLayer& L = layers_[l];
    const std::string b = S(kLlm, "layers." + std::to_string(l));
    const HostTensor *qw, *qb, *kw, *kb, *vw, *vb, *ow, *gw, *uw, *dw, *n1, *n2;
    if (int rc = need(b + ".input_layernorm.weight", &n1, {H})) return rc;
    if (int rc = need(b + ".post_attention_layernorm.weight", &n2, {H})) return rc;
    if (int rc = need(b + ".self_attn.q_proj.weight", &qw, {nq * hd, H})) return rc;
    if (int rc = need(b + ".self_attn.q_proj.bias", &qb, {nq * hd})) return rc;
    if (int rc = need(b + ".self_attn.k_proj.weight", &kw, {nkv * hd, H})) return rc;
    if (int rc = need(b + ".self_attn.k_proj.bias", &kb, {nkv * hd})) return rc;
    if (int rc = need(b + ".self_attn.v_proj.weight", &vw, {nkv * hd, H})) return rc;
    if (int rc = need(b + ".self_attn.v_proj.bias", &vb, {nkv * hd})) return rc;
    if (int rc = need(b + ".self_attn.o_proj.weight", &ow, {H, nq * hd})) return rc;
    if (int rc = need(b + ".mlp.gate_proj.weight", &gw, {I, H})) return rc;
    if (int rc = need(b + ".mlp.up_proj.weight", &uw, {I, H})) return rc;
    if (int rc = need(b + ".mlp.down_proj.weight", &dw, {H, I})) return rc;
    L.ln1 = upload_f32(n1->data);
    L.ln2 = upload_f32(n2->data);
    FVLA_REQUIRE(L.ln1 && L.ln2, "cudaMalloc failed (norm)");
    std::vector<float> w(qw->data);
    w.insert(w.end(), kw->data.begin(), kw->data.end());
    w.insert(w.end(), vw->data.begin(), vw->data.end());
    std::vector<float> bias(qb->data);
    bias.insert(bias.end(), kb->data.begin(), kb->data.end());
    bias.insert(bias.end(), vb->data.begin(), vb->data.end());
    if (int rc = make_gemm(&L.qkv, w, (nq + 2 * nkv) * hd, H, &bias)) return rc;
    if (int rc = make_gemm(&L.o, ow->data, H, nq * hd, nullptr)) return rc;
    std::vector<float> gu(static_cast<size_t>(2) * I * H);
    for (int j = 0; j < I; ++j) {
      std::memcpy(&gu[static_cast<size_t>(2 * j) * H], &gw->data[static_cast<size_t>(j) * H], sizeof(float) * H);
      std::memcpy(&gu[static_cast<size_t>(2 * j + 1) * H], &uw->data[static_cast<size_t>(j) * H], sizeof(float) * H);
    }
    if (int rc = make_gemm(&L.gate_up, gu, 2 * I, H, nullptr)) return rc;
    if (int rc = make_gemm(&L.down, dw->data, H, I, nullptr)) return rc;
  }
  if (int rc = need(S(kLlm, "norm.weight"), &t, {H})) return rc;
  final_norm_ = upload_f32(t->data);
  FVLA_REQUIRE(final_norm_ != nullptr, "cudaMalloc failed (final norm)");
  return 0;
}

int Engine::pack_head() {
  const int H = cfg.hidden, Sd = cfg.state_dim, Hd = cfg.hidden_dim, F = cfg.fusion_dim, A = cfg.action_dim;
  head_.H = H; head_.S = Sd; head_.Hd = Hd; head_.F = F; head_.A = A;
  // name -> (destination slot, element count, stored in engine dtype?)
  struct Slot { const char* name; const void** dst; int64_t n; bool act; };
  const Slot slots[] = {
      {"state_projection.0.weight", reinterpret_cast<const void**>(&head_.ln_s_w), Sd, false},
      {"state_projection.0.bias", reinterpret_cast<const void**>(&head_.ln_s_b), Sd, false},
      {"state_projection.1.weight", &head_.w_state, static_cast<int64_t>(Hd) * Sd, true},
      {"state_projection.1.bias", reinterpret_cast<const void**>(&head_.b_state), Hd, false},
      {"fusion.0.weight", &head_.w_f0, static_cast<int64_t>(F) * (H + Hd), true},
      {"fusion.0.bias", reinterpret_cast<const void**>(&head_.b_f0), F, false},
      {"fusion.1.weight", reinterpret_cast<const void**>(&head_.ln_f_w), F, false},
      {"fusion.1.bias", reinterpret_cast<const void**>(&head_.ln_f_b), F, false},
      {"fusion.4.weight", &head_.w_f4, static_cast<int64_t>(F) * F, true},
      {"fusion.4.bias", reinterpret_cast<const void**>(&head_.b_f4), F, false},
      {"action_head.weight", &head_.w_act, static_cast<int64_t>(A) * F, true},
      {"action_head.bias", reinterpret_cast<const void**>(&head_.b_act), A, false},
  };
  {
    // LeRobot (un)normaliser fused into the head: identity until fvla_set_io_normalization
    std::vector<float> ident(static_cast<size_t>(2 * Sd + 2 * A), 0.f);
    for (int i = 0; i < Sd; ++i) ident[Sd + i] = 1.f;
    for (int i = 0; i < A; ++i) ident[2 * Sd + i] = 1.f;
    io_norm_ = upload_f32(ident);
    FVLA_REQUIRE(io_norm_ != nullptr, "cudaMalloc failed (io normalisation)");
    head_.st_mean = io_norm_; head_.st_inv_std = io_norm_ + Sd;
    head_.act_scale = io_norm_ + 2 * Sd; head_.act_shift = io_norm_ + 2 * Sd + A;
  }
  for (const Slot& sl : slots) {
    const HostTensor* t;
    if (int rc = need(sl.name, &t, {sl.n})) return rc;
    void* p = sl.act ? upload_act(t->data) : static_cast<void*>(upload_f32(t->data));
    FVLA_REQUIRE(p != nullptr, "cudaMalloc failed (head)");
    *sl.dst = p;
    head_slots_[sl.name] = {p, sl.n, sl.act};
  }
  return 0;
}

// Trainable head parameters may change between forwards (FastVLAPolicy.forward trains only the
// head): overwrite the packed copy in place.
int Engine::update_head_tensor(const std::string& name, const HostTensor& t) {
  auto it = head_slots_.find(name);
  if (it == head_slots_.end()) {
    set_error("load_tensor after finalize is only allowed for action-head tensors, got: " + name);
    return 2;
  }
  FVLA_REQUIRE(t.numel() == it->second.n, "head tensor size changed");
  if (it->second.act && cfg.dtype == FVLA_BF16) {
    std::vector<__nv_bfloat16> h(t.data.size());
    for (size_t i = 0; i < h.size(); ++i) h[i] = __float2bfloat16_rn(t.data[i]);
    FVLA_CUDA_CHECK(cudaMemcpy(it->second.ptr, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  } else {
    FVLA_CUDA_CHECK(cudaMemcpy(it->second.ptr, t.data.data(), t.data.size() * 4, cudaMemcpyHostToDevice));
  }
  return 0;
}

int Engine::set_io_normalization(const float* state_mean, const float* state_inv_std, const float* action_scale,
                                 const float* action_shift) {
  FVLA_REQUIRE(finalized_ && io_norm_ != nullptr, "set_io_normalization: engine has no finalized action head");
  const int Sd = cfg.state_dim, A = cfg.action_dim;
  std::vector<float> v(static_cast<size_t>(2 * Sd + 2 * A), 0.f);
  for (int i = 0; i < Sd; ++i) {
    v[i] = state_mean ? state_mean[i] : 0.f;
    v[Sd + i] = state_inv_std ? state_inv_std[i] : 1.f;
  }
  for (int i = 0; i < A; ++i) {
    v[2 * Sd + i] = action_scale ? action_scale[i] : 1.f;
    v[2 * Sd + A + i] = action_shift ? action_shift[i] : 0.f;
  }
  // in place: captured graphs keep pointing at the same buffer
  FVLA_CUDA_CHECK(cudaMemcpy(io_norm_, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
  return 0;
}

int Engine::finalize() {
  FVLA_REQUIRE(!finalized_, "finalize called twice");
  FVLA_REQUIRE(cfg.dtype == FVLA_F32 || cfg.dtype == FVLA_BF16, "engine dtype must be fp32 or bf16");
  FVLA_REQUIRE(cfg.vis_num_stages >= 1 && cfg.vis_num_stages <= FVLA_MAX_VIS_STAGES, "vis_num_stages");
  int div = 4;
  for (int i = 0; i + 1 < cfg.vis_num_stages; ++i) div *= 2;
  FVLA_REQUIRE(cfg.image_size % div == 0, "image_size must be divisible by the tower's total stride");
  FVLA_REQUIRE(cfg.hidden % 8 == 0 && cfg.head_dim % 8 == 0 && cfg.intermediate % 8 == 0, "Qwen2 dims % 8");
  std::vector<std::string> miss;
  missing(&miss);
  if (!miss.empty()) {
    set_error("finalize: " + std::to_string(miss.size()) + " tensors missing, first: " + miss[0]);
    return 3;
  }
  if (int rc = pack_vision()) return rc;
  if (int rc = pack_decoder()) return rc;
  if (cfg.state_dim > 0)
    if (int rc = pack_head()) return rc;
  host_.clear();
  finalized_ = true;
  FVLA_CUDA_CHECK(cudaDeviceSynchronize());
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Workspace
// ---------------------------------------------------------------------------------------------
int Engine::ensure(const std::string& name, size_t bytes, void** out) {
  bytes = (bytes + 255) & ~static_cast<size_t>(255);
  auto it = ws_.bufs.find(name);
  if (it != ws_.bufs.end() && it->second.second >= bytes) { *out = it->second.first; return 0; }
  if (it != ws_.bufs.end()) {
    FVLA_CUDA_CHECK(cudaDeviceSynchronize());
    FVLA_CUDA_CHECK(cudaFree(it->second.first));
    ws_.total -= it->second.second;
    ws_.bufs.erase(it);
  }
  void* p = nullptr;
  FVLA_CUDA_CHECK(cudaMalloc(&p, bytes));
  clear_graphs();  // captured launches hold workspace pointers
  ws_.bufs[name] = {p, bytes};
  ws_.total += bytes;
  *out = p;
  return 0;
}

int Engine::set_tap(int stage, void* dst, int64_t cap) {
  if (dst == nullptr) taps_.erase(stage);
  else taps_[stage] = {dst, cap};
  return 0;
}

int Engine::tap(int stage, const void* src, size_t bytes, size_t off, cudaStream_t s) {
  auto it = taps_.find(stage);
  if (it == taps_.end()) return 0;
  if (static_cast<int64_t>(off + bytes) > it->second.second) {
    set_error("tap buffer for stage " + std::to_string(stage) + " too small: need " +
              std::to_string(off + bytes) + " bytes");
    return 2;
  }
  FVLA_CUDA_CHECK(cudaMemcpyAsync(static_cast<char*>(it->second.first) + off, src, bytes,
                                  cudaMemcpyDeviceToDevice, s));
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Forward
// ---------------------------------------------------------------------------------------------
int Engine::run_gemm(const GemmW& w, const void* A, void* D, int M, int act, const void* resid,
                     bool swiglu, cudaStream_t s, bool ab_f16, bool out_f32, bool rope, int split_k, int block_n) {
  GemmArgs g;
  g.block_n = block_n;
  g.A = A; g.lda = w.K;
  g.W = w.w; g.ldw = w.K;
  g.D = D; g.ldd = swiglu ? w.N / 2 : w.N;
  g.M = M; g.N = w.N; g.K = w.K;
  g.bias = w.bias;
  g.resid = resid; g.ldr = w.N;
  g.act = (act == ACT_GELU && w.half_in) ? static_cast<int>(ACT_GELU_HALF) : act;
  g.swiglu = swiglu ? 1 : 0;
  g.ab_f16 = ab_f16 ? 1 : 0;
  g.out_f32 = (out_f32 && cfg.dtype == FVLA_BF16) ? 1 : 0;  // the fp32 mode's D is FP32 anyway
  if (cfg.dtype == FVLA_BF16) {
    if (rope) {
      g.rope_tab = rope_tab_; g.rope_T = merged_T_;
      g.rope_cols = (cfg.n_q_heads + cfg.n_kv_heads) * cfg.head_dim;
    }
    if (g.out_f32 && resid == D) g.split_k = split_k;
  }
  ++launches;
  const double fl = 2.0 * M * static_cast<double>(w.N) * w.K;
  flops += fl * flop_scale_;
  prof_begin(s);
  const int rc = gemm(cfg.dtype, g, s);
  if (profile_) {
    const double e = static_cast<double>(esz());
    const double by = e * (static_cast<double>(M) * w.K + static_cast<double>(w.N) * w.K +
                           static_cast<double>(M) * g.ldd + (resid ? static_cast<double>(M) * w.N : 0.0));
    prof_end(std::string(prof_scope_) + "gemm M" + std::to_string(M) + " N" + std::to_string(w.N) + " K" +
                 std::to_string(w.K) + (swiglu ? " swiglu" : "") + (resid ? " +res" : "") +
                 (g.rope_tab ? " rope" : "") + (g.split_k > 1 ? " splitk" + std::to_string(g.split_k) : "") +
                 ((act == ACT_GELU || act == ACT_GELU_HALF_F16) ? " gelu" : ""),
             fl, by, s);
  }
  return rc;
}

// ConvFFN tail: out = resid + fc2(gelu(fc1(z))).  Fused on chip where the kernel covers the width (stages 0/1 in
// bf16), otherwise two GEMMs through the 4x hidden buffer `hid`.
int Engine::run_ffn(const VisBlock& blk, const void* z, void* hid, void* out_resid, int M, cudaStream_t s) {
  const int d = blk.fc1.K, hd = blk.fc1.N;
  static const bool fuse_ffn = std::getenv("FVLA_DISABLE_FUSED_FFN") == nullptr;  // A/B switch for profiling
  if (fuse_ffn && blk.fc1.half_in && blk.fc1.bias != nullptr && blk.fc2.bias != nullptr && blk.fc2.w_f16 != nullptr &&
      ffn_fused_supported(cfg.dtype, d, hd)) {
    FfnFusedArgs a;
    a.x = z; a.w1 = blk.fc1.w; a.b1 = blk.fc1.bias; a.w2 = blk.fc2.w_f16; a.b2 = blk.fc2.bias;
    a.resid = out_resid; a.out = out_resid; a.M = M; a.C = d; a.hidden = hd;
    ++launches;
    const double fl = 4.0 * M * static_cast<double>(hd) * d;
    flops += fl;
    prof_begin(s);
    const int rc = ffn_fused(a, s);
    prof_end("vis.ffn_fused M" + std::to_string(M) + " C" + std::to_string(d) + " H" + std::to_string(hd), fl,
             static_cast<double>(esz()) * (3.0 * M * d + 2.0 * hd * d), s);
    return rc;
  }
  if (cfg.dtype == FVLA_BF16 && blk.fc1.half_in && blk.fc2.w_f16 != nullptr) {
    if (int rc = run_gemm(blk.fc1, z, hid, M, ACT_GELU_HALF_F16, nullptr, false, s)) return rc;
    GemmW fc2h = blk.fc2;
    fc2h.w = blk.fc2.w_f16;
    return run_gemm(fc2h, hid, out_resid, M, ACT_NONE, out_resid, false, s, /*ab_f16=*/true);
  }
  if (int rc = run_gemm(blk.fc1, z, hid, M, ACT_GELU, nullptr, false, s)) return rc;
  return run_gemm(blk.fc2, hid, out_resid, M, ACT_NONE, out_resid, false, s);
}

int Engine::run_dw(const DwW& w, const void* in, void* out, int B, int H, int W, cudaStream_t s) {
  ++launches;
  const int Ho = (H - 1) / w.stride + 1, Wo = (W - 1) / w.stride + 1;
  const double fl = 2.0 * w.k * w.k * static_cast<double>(B) * Ho * Wo * w.cin * w.mult;
  flops += fl;
  prof_begin(s);
  const int rc = dwconv(cfg.dtype, in, w.w, w.bias, out, B, H, W, w.cin, w.mult, w.k, w.stride, w.act, s, w.wtab);
  if (profile_) {
    const double e = static_cast<double>(esz());
    const double by = e * (static_cast<double>(B) * H * W * w.cin + static_cast<double>(B) * Ho * Wo * w.cin * w.mult) +
                      4.0 * (w.k * w.k + 1) * w.cin * w.mult;
    prof_end("vis.dwconv k" + std::to_string(w.k) + " s" + std::to_string(w.stride) + " m" +
                 std::to_string(w.mult) + " C" + std::to_string(w.cin) + " HW" + std::to_string(H),
             fl, by, s);
  }
  return rc;
}

int Engine::reserve(int B, int n_tokens) {
  FVLA_REQUIRE(finalized_, "reserve before finalize");
  FVLA_REQUIRE(B > 0 && n_tokens > 0, "reserve: empty");
  const size_t e = esz();
  const int S = cfg.image_size;
  int chunk = cfg.vision_chunk > 0 ? cfg.vision_chunk : 32;
  chunk = std::min(chunk, B);
  // per-image maxima over the tower
  size_t act_max = 0, hid_max = 0;
  int side = S / 4;
  for (int i = 0; i < cfg.vis_num_stages; ++i) {
    const size_t px = static_cast<size_t>(side) * side;
    const int d = cfg.vis_dims[i];
    act_max = std::max(act_max, px * d);
    hid_max = std::max(hid_max, px * d * std::max(cfg.vis_mlp_ratio, 3));
    if (i + 1 < cfg.vis_num_stages) side /= 2;
  }
  act_max = std::max(act_max, static_cast<size_t>(side) * side * mm_hidden());
  hid_max = std::max(hid_max, static_cast<size_t>(S / 2) * (S / 2) * cfg.vis_dims[0]);  // stem.0 output
  void* p;
  if (int rc = ensure("vis_pre", static_cast<size_t>(chunk) * S * S * 4 * e, &p)) return rc;
  if (cfg.dtype == FVLA_BF16)
    if (int rc = ensure("vis_col", static_cast<size_t>(chunk) * (S / 2) * (S / 2) * 32 * e, &p)) return rc;
  if (int rc = ensure("vis_x", chunk * act_max * e, &p)) return rc;
  if (int rc = ensure("vis_y", chunk * act_max * e, &p)) return rc;
  if (int rc = ensure("vis_z", chunk * act_max * e, &p)) return rc;
  if (int rc = ensure("vis_h", chunk * hid_max * e, &p)) return rc;
  if (int rc = ensure("se_mean", static_cast<size_t>(B) * mm_hidden() * 4, &p)) return rc;
  if (int rc = ensure("se_gate", static_cast<size_t>(B) * mm_hidden() * 4, &p)) return rc;
  const int nimg = n_img_tokens();
  const int H = cfg.hidden;
  if (int rc = ensure("feats", static_cast<size_t>(B) * nimg * mm_hidden() * e, &p)) return rc;
  if (int rc = ensure("proj_h", static_cast<size_t>(B) * nimg * H * e, &p)) return rc;
  if (int rc = ensure("img_tok", static_cast<size_t>(B) * nimg * H * e, &p)) return rc;
  const size_t Tm = static_cast<size_t>(n_tokens) + nimg;  // upper bound of the merged length
  const size_t rows = static_cast<size_t>(B) * Tm;
  const int qkv_n = (cfg.n_q_heads + 2 * cfg.n_kv_heads) * cfg.head_dim;
  if (int rc = ensure("dec_x", rows * H * 4, &p)) return rc;  // residual stream: FP32 in either mode
  if (int rc = ensure("dec_xn", rows * H * e, &p)) return rc;
  if (int rc = ensure("dec_qkv", rows * qkv_n * e, &p)) return rc;
  if (int rc = ensure("dec_ao", rows * cfg.n_q_heads * cfg.head_dim * e, &p)) return rc;
  if (int rc = ensure("dec_act", rows * cfg.intermediate * e, &p)) return rc;
  if (int rc = ensure("plan", rows * 4, &p)) return rc;
  if (int rc = ensure("pool_idx", static_cast<size_t>(B) * 4, &p)) return rc;
  if (int rc = ensure("lens", static_cast<size_t>(B) * 4, &p)) return rc;
  if (int rc = ensure("pooled", static_cast<size_t>(B) * H * 4, &p)) return rc;
  if (int rc = ensure("actions", static_cast<size_t>(B) * cfg.action_dim * 4, &p)) return rc;
  if (int rc = ensure("tap_state", static_cast<size_t>(B) * cfg.hidden_dim * 4, &p)) return rc;
  if (int rc = ensure("tap_fused", static_cast<size_t>(B) * cfg.fusion_dim * 4, &p)) return rc;
  if (int rc = ensure("head_x1", static_cast<size_t>(B) * cfg.fusion_dim * 4, &p)) return rc;
  // rotary tables for every merged position (HF default rope init, fp32)
  if (rope_len_ < static_cast<int>(Tm)) {
    const int half = cfg.head_dim / 2;
    std::vector<float> c(Tm * half), sn(Tm * half);
    for (size_t pos = 0; pos < Tm; ++pos)
      for (int i = 0; i < half; ++i) {
        const float inv_freq =
            1.0f / std::pow(cfg.rope_theta, static_cast<float>(2 * i) / static_cast<float>(cfg.head_dim));
        const float ang = static_cast<float>(pos) * inv_freq;
        c[pos * half + i] = std::cos(ang);
        sn[pos * half + i] = std::sin(ang);
      }
    void *pc, *ps;
    if (int rc = ensure("rope_cos", c.size() * 4, &pc)) return rc;
    if (int rc = ensure("rope_sin", sn.size() * 4, &ps)) return rc;
    FVLA_CUDA_CHECK(cudaMemcpy(pc, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
    FVLA_CUDA_CHECK(cudaMemcpy(ps, sn.data(), sn.size() * 4, cudaMemcpyHostToDevice));
    {
      std::vector<uint32_t> tab(c.size());
      for (size_t i = 0; i < c.size(); ++i) {
        const __half2 h = __floats2half2_rn(c[i], sn[i]);
        std::memcpy(&tab[i], &h, 4);
      }
      void* pt;
      if (int rc = ensure("rope_tab", tab.size() * 4, &pt)) return rc;
      FVLA_CUDA_CHECK(cudaMemcpy(pt, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice));
      rope_tab_ = static_cast<uint32_t*>(pt);
    }
    rope_cos_ = static_cast<float*>(pc);
    rope_sin_ = static_cast<float*>(ps);
    rope_len_ = static_cast<int>(Tm);
  }
  return 0;
}

int Engine::vision_chunk(const fvla_forward_args& a, int c0, int bc, void* feats, cudaStream_t s) {
  const size_t e = esz();
  const int S = cfg.image_size;
  char* pre = static_cast<char*>(ws_.bufs["vis_pre"].first);
  char* X = static_cast<char*>(ws_.bufs["vis_x"].first);
  char* Y = static_cast<char*>(ws_.bufs["vis_y"].first);
  char* Z = static_cast<char*>(ws_.bufs["vis_z"].first);
  char* Hb = static_cast<char*>(ws_.bufs["vis_h"].first);

  // ---- ingest ----
  PreprocessArgs pa;
  const size_t img_elems = static_cast<size_t>(a.img_c) * a.img_h * a.img_w;
  pa.src = static_cast<const char*>(a.images) + c0 * img_elems * dtype_size(a.img_dtype);
  pa.src_dtype = a.img_dtype; pa.src_nhwc = a.img_nhwc;
  pa.B = bc; pa.C = a.img_c; pa.h = a.img_h; pa.w = a.img_w; pa.S = S;
  pa.letterbox = a.letterbox; pa.pad_value = a.pad_value;
  pa.scale = a.img_scale; pa.normalize = a.normalize;
  for (int i = 0; i < 3; ++i) { pa.mean[i] = a.mean[i]; pa.inv_std[i] = a.inv_std[i]; }
  pa.dst = pre;
  ++launches;
  prof_scope_ = "vis.";
  prof_begin(s);
  if (int rc = preprocess_images(cfg.dtype, pa, s)) return rc;
  prof_end("vis.preprocess", 0.0,
           static_cast<double>(bc) * (img_elems * dtype_size(a.img_dtype) + static_cast<double>(S) * S * 4 * e), s);
  if (int rc = tap(FVLA_TAP_PREPROCESS, pre, static_cast<size_t>(bc) * S * S * 4 * e,
                   static_cast<size_t>(c0) * S * S * 4 * e, s)) return rc;

  // ---- stem ----
  const int d0 = cfg.vis_dims[0];
  static const bool fuse_stem = std::getenv("FVLA_DISABLE_FUSED_STEM") == nullptr;  // A/B switch for profiling
  if (fuse_stem && stem_btab_ != nullptr) {
    // stem.0 (tensor cores, implicit GEMM) + GELU + stem.1 (depthwise 3x3 s2) + GELU in one kernel
    ++launches;
    const double fl = 2.0 * 27 * static_cast<double>(bc) * (S / 2) * (S / 2) * d0 +
                      2.0 * 9 * static_cast<double>(bc) * (S / 4) * (S / 4) * d0;
    flops += fl;
    prof_begin(s);
    if (int rc = stem_fused(pre, stem_btab_, stem0_bh_, stem1_.w, stem1_.bias, Z, bc, S, d0, s)) return rc;
    prof_end("vis.stem_fused", fl,
             static_cast<double>(bc) * e * (static_cast<double>(S) * S * 4 + static_cast<double>(S / 4) * (S / 4) * d0), s);
  } else {
    if (cfg.dtype == FVLA_BF16) {
      // 27-tap patches -> [M, 32] bf16, then a K=32 tensor-core GEMM with bias + GELU in the epilogue
      char* col = static_cast<char*>(ws_.bufs["vis_col"].first);
      const int Ms = bc * (S / 2) * (S / 2);
      ++launches;
      prof_begin(s);
      if (int rc = stem_im2col_bf16(pre, col, bc, S, S, s)) return rc;
      prof_end("vis.stem_im2col", 0.0, static_cast<double>(bc) * e * (static_cast<double>(S) * S * 4 + static_cast<double>(Ms / bc) * 32), s);
      const double before = flops;
      if (int rc = run_gemm(stem0_gemm_, col, Hb, Ms, ACT_GELU, nullptr, false, s)) return rc;
      flops = before + 2.0 * 27 * static_cast<double>(Ms) * d0;  // algorithmic taps, not the zero padding
    } else {
      ++launches;
      flops += 2.0 * 27 * static_cast<double>(bc) * (S / 2) * (S / 2) * d0;
      prof_begin(s);
      if (int rc = stem_conv3x3_s2(cfg.dtype, pre, stem0_w_, stem0_b_, Hb, bc, S, S, d0, s)) return rc;
      prof_end("vis.stem_conv3x3", 2.0 * 27 * static_cast<double>(bc) * (S / 2) * (S / 2) * d0,
               static_cast<double>(bc) * e * (static_cast<double>(S) * S * 4 + static_cast<double>(S / 2) * (S / 2) * d0), s);
    }
    if (int rc = run_dw(stem1_, Hb, Z, bc, S / 2, S / 2, s)) return rc;
  }
  int side = S / 4;
  if (int rc = run_gemm(stem2_, Z, X, bc * side * side, ACT_GELU, nullptr, false, s)) return rc;
  if (int rc = tap(FVLA_TAP_STEM, X, static_cast<size_t>(bc) * side * side * d0 * e,
                   static_cast<size_t>(c0) * side * side * d0 * e, s)) return rc;

  // ---- stages ----
  for (int i = 0; i < cfg.vis_num_stages; ++i) {
    VisStage& st = stages_[i];
    const int d = st.dim;
    const int M = bc * side * side;
    if (st.has_cpe) {
      if (int rc = run_dw(st.cpe, X, Y, bc, side, side, s)) return rc;
      std::swap(X, Y);
    }
    for (auto& blk : st.blocks) {
      if (!blk.attn) {
        // RepMixerBlock: x = mixer(x); x = x + ls * ConvFFN(x)
        if (int rc = run_dw(blk.mixer, X, Y, bc, side, side, s)) return rc;
        if (int rc = run_dw(blk.ffn_dw, Y, Z, bc, side, side, s)) return rc;
        if (int rc = run_ffn(blk, Z, Hb, Y, M, s)) return rc;
        std::swap(X, Y);
      } else {
        // AttentionBlock: x = x + ls1 * MHSA(BN(x)); x = x + ls2 * ConvFFN(x)
        const char* qkv_in = X;
        if (blk.attn_layernorm) {
          ++launches;
          prof_begin(s);
          if (int rc = layernorm_rows(cfg.dtype, X, Y, M, d, kLnChannelEps, s)) return rc;
          prof_end("vis.layernorm C" + std::to_string(d), 0.0, 2.0 * M * static_cast<double>(d) * e, s);
          qkv_in = Y;
        }
        if (int rc = run_gemm(blk.qkv, qkv_in, Hb, M, ACT_NONE, nullptr, false, s)) return rc;
        AttnArgs at;
        at.q = Hb; at.k = Hb + static_cast<size_t>(d) * e; at.v = Hb + static_cast<size_t>(2 * d) * e;
        at.ld_qkv = 3 * d; at.o = Z; at.ld_o = d;
        at.B = bc; at.N = side * side;
        at.heads_q = at.heads_kv = d / cfg.vis_head_dim; at.head_dim = cfg.vis_head_dim;
        at.scale = 1.0f / std::sqrt(static_cast<float>(cfg.vis_head_dim));
        at.causal = 0;
        ++launches;
        flops += 4.0 * bc * static_cast<double>(at.N) * at.N * d;
        prof_begin(s);
        if (int rc = attention(cfg.dtype, at, s)) return rc;
        prof_end("vis.attention N" + std::to_string(at.N) + " h" + std::to_string(at.heads_q),
                 4.0 * bc * static_cast<double>(at.N) * at.N * d, 4.0 * M * static_cast<double>(d) * e, s);
        if (int rc = run_gemm(blk.proj, Z, X, M, ACT_NONE, X, false, s)) return rc;
        if (int rc = run_dw(blk.ffn_dw, X, Z, bc, side, side, s)) return rc;
        if (int rc = run_ffn(blk, Z, Hb, X, M, s)) return rc;
      }
    }
    if (int rc = tap(FVLA_TAP_VIS_STAGE0 + i, X, static_cast<size_t>(M) * d * e,
                     static_cast<size_t>(c0) * side * side * d * e, s)) return rc;
    if (st.has_down) {
      if (int rc = run_dw(st.down_dw, X, Z, bc, side, side, s)) return rc;
      side /= 2;
      if (int rc = run_gemm(st.down_pw, Z, X, bc * side * side, ACT_GELU, nullptr, false, s)) return rc;
    }
  }
  // ---- conv_exp (grouped 3x3, x2 channels) -> pre-SE image features of this chunk ----
  const int ce = mm_hidden(), hw = side * side;
  char* dst = static_cast<char*>(feats) + static_cast<size_t>(c0) * hw * ce * e;
  if (int rc = run_dw(exp_dw_, X, dst, bc, side, side, s)) return rc;
  return 0;
}

int Engine::forward(const fvla_forward_args& a, cudaStream_t s) {
  FVLA_REQUIRE(finalized_, "forward before finalize");
  FVLA_REQUIRE(a.batch > 0 && a.n_tokens > 0, "forward: empty batch");
  FVLA_REQUIRE(a.token_ids != nullptr && a.text_len != nullptr, "forward: token_ids/text_len required");
  const int B = a.batch, T = a.n_tokens, nimg = n_img_tokens();
  const size_t e = esz();
  launches = 0;
  flops = 0.0;
  if (int rc = reserve(B, T)) return rc;

  // ---- LLaVA splice plan (prepare_inputs_labels_for_multimodal [EXT]) ----
  std::vector<int> mlen(B);
  bool any_image = false;
  int Tm = 1;
  for (int b = 0; b < B; ++b) {
    const int tl = a.text_len[b];
    FVLA_REQUIRE(tl >= 0 && tl <= T, "text_len out of range");
    int nph = 0;
    for (int t = 0; t < tl; ++t) nph += a.token_ids[static_cast<size_t>(b) * T + t] == FVLA_IMAGE_TOKEN_INDEX;
    FVLA_REQUIRE(nph <= 1, "at most one image placeholder per sample is supported");
    any_image = any_image || nph > 0;
    mlen[b] = tl - nph + nph * nimg;
    Tm = std::max(Tm, mlen[b]);
  }
  merged_len = Tm;
  {
    double rows = 0.0, sq = 0.0;
    for (int b = 0; b < B; ++b) { rows += mlen[b]; sq += static_cast<double>(mlen[b]) * mlen[b]; }
    dec_row_frac_ = Tm > 0 ? rows / (static_cast<double>(B) * Tm) : 1.0;
    dec_sq_frac_ = Tm > 0 ? sq / (static_cast<double>(B) * Tm * Tm) : 1.0;
  }
  // Host staging of the splice plan: a ring of PINNED buffers, each guarded by an event, so the upload is a true
  // asynchronous copy (a pageable source above the driver's inline limit makes cudaMemcpyAsync stage + wait on the
  // host, which would serialise the caller's launch-ahead with the GPU).
  const size_t n_plan = static_cast<size_t>(B) * Tm, n_ints = n_plan + 2 * static_cast<size_t>(B);
  HostStage& hs = host_ring_[host_ring_next_];
  host_ring_next_ = (host_ring_next_ + 1) % kHostRing;
  if (hs.ev == nullptr) FVLA_CUDA_CHECK(cudaEventCreateWithFlags(&hs.ev, cudaEventDisableTiming));
  else FVLA_CUDA_CHECK(cudaEventSynchronize(hs.ev));  // the copy that last read this slot has completed
  if (hs.cap < n_ints) {
    if (hs.p != nullptr) FVLA_CUDA_CHECK(cudaFreeHost(hs.p));
    hs.p = nullptr; hs.cap = 0;
    FVLA_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&hs.p), n_ints * 4 + 64));
    hs.cap = n_ints;
  }
  int* plan = hs.p; int* pidx = hs.p + n_plan; int* lens = pidx + B;
  std::fill(plan, plan + n_plan, -1);
  for (int b = 0; b < B; ++b) {
    int pos = 0;
    for (int t = 0; t < a.text_len[b]; ++t) {
      const int id = a.token_ids[static_cast<size_t>(b) * T + t];
      if (id == FVLA_IMAGE_TOKEN_INDEX) {
        for (int j = 0; j < nimg; ++j) plan[static_cast<size_t>(b) * Tm + pos++] = -2 - j;
      } else {
        FVLA_REQUIRE(id >= 0 && id < cfg.vocab, "token id out of vocabulary");
        plan[static_cast<size_t>(b) * Tm + pos++] = id;
      }
    }
    // reference-literal pooling index: text attention-mask length - 1 (fastvlm_adapter.py:353-358)
    int pi = a.pool_idx != nullptr ? a.pool_idx[b] : a.text_len[b] - 1;
    pidx[b] = std::min(std::max(pi, 0), Tm - 1);
    lens[b] = std::min(a.text_len[b], Tm);
  }
  int* d_plan = static_cast<int*>(ws_.bufs["plan"].first);
  int* d_pidx = static_cast<int*>(ws_.bufs["pool_idx"].first);
  int* d_lens = static_cast<int*>(ws_.bufs["lens"].first);
  FVLA_CUDA_CHECK(cudaMemcpyAsync(d_plan, plan, n_plan * 4, cudaMemcpyHostToDevice, s));
  FVLA_CUDA_CHECK(cudaMemcpyAsync(d_pidx, pidx, static_cast<size_t>(B) * 4, cudaMemcpyHostToDevice, s));
  FVLA_CUDA_CHECK(cudaMemcpyAsync(d_lens, lens, static_cast<size_t>(B) * 4, cudaMemcpyHostToDevice, s));
  FVLA_CUDA_CHECK(cudaEventRecord(hs.ev, s));

  // ---- CUDA-graph replay for small batches (the b=1 select_action latency is launch-bound: ~390 launches) ----
  static const bool graphs_on = std::getenv("FVLA_DISABLE_GRAPHS") == nullptr;
  const bool graphable = graphs_on && B <= kGraphMaxBatch && !profile_ && taps_.empty() && a.pooled == nullptr &&
                         a.states != nullptr && a.actions != nullptr && (a.images != nullptr || !any_image);
  if (!graphable) return launch_all(a, B, Tm, any_image, s);
  return forward_graph(a, B, Tm, any_image, s);
}

// Kernel sequence of one forward; reads the splice plan / pool indices from the workspace (uploaded by forward()).
// Pure stream-ordered launches (no allocation, no synchronisation): capturable into a CUDA graph.
int Engine::launch_all(const fvla_forward_args& a, int B, int Tm, bool any_image, cudaStream_t s) {
  flop_scale_ = 1.0;
  const int H = cfg.hidden, nimg = n_img_tokens();
  const size_t e = esz();
  int* d_plan = static_cast<int*>(ws_.bufs["plan"].first);
  int* d_pidx = static_cast<int*>(ws_.bufs["pool_idx"].first);
  int* d_lens = static_cast<int*>(ws_.bufs["lens"].first);
  // ---- FastViTHD + projector ----
  void* feats = ws_.bufs["feats"].first;
  void* proj_h = ws_.bufs["proj_h"].first;
  void* img_tok = ws_.bufs["img_tok"].first;
  if (any_image || !cfg.skip_unused_vision) {
    FVLA_REQUIRE(a.images != nullptr, "forward: images required");
    int chunk = cfg.vision_chunk > 0 ? cfg.vision_chunk : 32;
    chunk = std::min(chunk, B);
    for (int c0 = 0; c0 < B; c0 += chunk) {
      const int bc = std::min(chunk, B - c0);
      if (int rc = vision_chunk(a, c0, bc, feats, s)) return rc;
    }
    // squeeze-excite + GELU over the whole batch, in place (conv_exp tail)
    {
      const int ce = mm_hidden();
      launches += 3;
      prof_begin(s);
      if (int rc = se_gelu(cfg.dtype, feats, feats, B, nimg, ce, cfg.vis_se_reduced, se_w1_, se_b1_, se_w2_,
                           se_b2_, static_cast<float*>(ws_.bufs["se_mean"].first),
                           static_cast<float*>(ws_.bufs["se_gate"].first), s)) return rc;
      prof_end("vis.se_gelu", 0.0, 3.0 * B * static_cast<double>(nimg) * ce * e, s);
    }
    if (int rc = tap(FVLA_TAP_IMAGE_FEATURES, feats, static_cast<size_t>(B) * nimg * mm_hidden() * e, 0, s)) return rc;
    prof_scope_ = "proj.";
    if (int rc = run_gemm(proj0_, feats, proj_h, B * nimg, ACT_GELU, nullptr, false, s)) return rc;
    if (int rc = run_gemm(proj2_, proj_h, img_tok, B * nimg, ACT_NONE, nullptr, false, s)) return rc;
    if (int rc = tap(FVLA_TAP_PROJECTOR, img_tok, static_cast<size_t>(B) * nimg * H * e, 0, s)) return rc;
  }

  // ---- Qwen2 prefill ----
  char* X = static_cast<char*>(ws_.bufs["dec_x"].first);
  char* Xn = static_cast<char*>(ws_.bufs["dec_xn"].first);
  char* QKV = static_cast<char*>(ws_.bufs["dec_qkv"].first);
  char* AO = static_cast<char*>(ws_.bufs["dec_ao"].first);
  char* ACTB = static_cast<char*>(ws_.bufs["dec_act"].first);
  const int M = B * Tm;
  ++launches;
  prof_scope_ = "llm.";
  prof_begin(s);
  // The residual stream X is FP32 in both precision modes: in bf16 mode every o-proj / down-proj epilogue adds
  // its fp32 accumulator to the fp32 stream (no rounding of the stream per layer), RMSNorm reads it and writes the
  // bf16 GEMM operand.  Taps of the stream are converted to the engine dtype.
  const int stream32 = cfg.dtype == FVLA_BF16 ? 1 : 0;
  auto tap_stream = [&](int stage) -> int {
    auto it = taps_.find(stage);
    if (it == taps_.end()) return 0;
    const size_t need = static_cast<size_t>(M) * H * e;
    if (static_cast<int64_t>(need) > it->second.second) {
      set_error("tap buffer for stage " + std::to_string(stage) + " too small: need " + std::to_string(need) + " bytes");
      return 2;
    }
    return convert(DT_F32, X, cfg.dtype, it->second.first, static_cast<long long>(M) * H, s);
  };
  if (int rc = embed_splice(cfg.dtype, embed_, img_tok, nimg, d_plan, X, B, Tm, H, s, stream32)) return rc;
  prof_end("llm.embed_splice", 0.0, M * static_cast<double>(H) * (e + 4), s);
  if (int rc = tap_stream(FVLA_TAP_EMBEDS)) return rc;
  const int nq = cfg.n_q_heads, nkv = cfg.n_kv_heads, hd = cfg.head_dim;
  merged_T_ = Tm;
  // bf16-mode layer fusions:
  //  * RoPE runs in the epilogue of the qkv projection (head_dim 64: a 64-column epilogue sub-tile is one head)
  //  * small batches (the b = 1 select_action prefill, M = T' rows): the 28 output tiles of the o-proj / down-proj
  //    cannot fill 74 CTA pairs while each walks the whole K loop, so those GEMMs split K over several pairs — their
  //    epilogue already reduce-adds into the fp32 stream, which makes the partial sums order-free
  // RMSNorm stays a kernel of its own: folding it into the projections (norm weight into W's columns, rstd as a row
  // scale of the accumulator) needs the producer's epilogue to emit the bf16 operand rows and their sums of squares,
  // i.e. to READ the fp32 stream it now only reduce-adds into.  Built and measured at batch 64: o-proj 1.13 -> 3.32,
  // down-proj 3.18 -> 4.24 ms/step against 0.98 ms/step for the 48 norm launches it removed (profiles/): the epilogue
  // then moves 2.5x the bytes of the K = 896 GEMM's operands.  The norm kernel runs at 4.6 TB/s.
  static const bool fuse_on = std::getenv("FVLA_DISABLE_LAYER_FUSION") == nullptr;  // A/B switch for profiling
  const bool rope_fused = fuse_on && cfg.dtype == FVLA_BF16 && hd == 64 && Tm >= 32 && rope_tab_ != nullptr;
  int split_o = 0, split_down = 0, bn_down = 0;
  if (fuse_on && cfg.dtype == FVLA_BF16) {
    const int pairs = num_sms() / 2;
    const int tiles = ceil_div(M, 256) * ceil_div(H, 64);   // small-M GEMMs run 64-wide tiles
    const int want = pairs / std::max(1, tiles);
    if (want >= 2) {
      split_o = std::min(want, std::max(1, (nq * hd) / 256));          // >= 4 k-blocks per split
      split_down = std::min(std::min(want, 8), std::max(1, cfg.intermediate / 256));
      // long K: a 64-wide tile moves 16 KB of A per 4 KB of W through L2 -> SM for every k-block; 128-wide tiles with
      // twice the splits keep the same number of busy pairs at 2/3 of the operand bytes (scripts/sweep_gemm_tiles.py,
      // M 272 x N 896 x K 4864: 64/2 12.9 us, 64/4 11.7, 128/4 9.3, 256/8 9.9)
      if (cfg.intermediate >= 2048 && H % 128 == 0) {
        bn_down = 128;
        const int tiles128 = ceil_div(M, 256) * ceil_div(H, 128);
        split_down = std::min(std::min(std::max(2, pairs / std::max(1, tiles128)), 8), std::max(1, cfg.intermediate / 256));
      }
    }
  }
  flop_scale_ = dec_row_frac_;
  for (int l = 0; l < cfg.n_layers; ++l) {
    DecLayer& L = layers_[l];
    ++launches;
    prof_begin(s);
    if (int rc = rmsnorm(cfg.dtype, X, L.ln1, Xn, M, H, cfg.rms_eps, s, stream32)) return rc;
    prof_end("llm.rmsnorm", 0.0, M * static_cast<double>(H) * (e + 4), s);
    if (int rc = run_gemm(L.qkv, Xn, QKV, M, ACT_NONE, nullptr, false, s, false, false, rope_fused)) return rc;
    AttnArgs at;
    at.q = QKV; at.k = QKV + static_cast<size_t>(nq * hd) * e; at.v = QKV + static_cast<size_t>((nq + nkv) * hd) * e;
    at.ld_qkv = (nq + 2 * nkv) * hd; at.o = AO; at.ld_o = nq * hd;
    at.B = B; at.N = Tm; at.heads_q = nq; at.heads_kv = nkv; at.head_dim = hd;
    at.scale = 1.0f / std::sqrt(static_cast<float>(hd));
    at.causal = 1;  // q and k arrive rotated (qkv epilogue, or in place just below)
    if (!rope_fused) {
      ++launches;
      prof_begin(s);
      if (int rc = rope_inplace(cfg.dtype, QKV, at.ld_qkv, B, Tm, nq + nkv, hd, rope_cos_, rope_sin_, s)) return rc;
      prof_end("llm.rope", 0.0, 2.0 * M * static_cast<double>((nq + nkv) * hd) * e, s);
    }
    ++launches;
    flops += 2.0 * B * static_cast<double>(Tm) * Tm * nq * hd * dec_sq_frac_;  // causal: half of 4*T^2*d, valid tokens
    prof_begin(s);
    if (int rc = attention(cfg.dtype, at, s)) return rc;
    prof_end("llm.attention T" + std::to_string(Tm), 2.0 * B * static_cast<double>(Tm) * Tm * nq * hd,
             static_cast<double>(M) * e * ((nq + 2 * nkv) * hd + nq * hd), s);
    if (int rc = run_gemm(L.o, AO, X, M, ACT_NONE, X, false, s, false, true, false, split_o)) return rc;
    ++launches;
    prof_begin(s);
    if (int rc = rmsnorm(cfg.dtype, X, L.ln2, Xn, M, H, cfg.rms_eps, s, stream32)) return rc;
    prof_end("llm.rmsnorm", 0.0, M * static_cast<double>(H) * (e + 4), s);
    if (int rc = run_gemm(L.gate_up, Xn, ACTB, M, ACT_NONE, nullptr, true, s)) return rc;
    if (int rc = run_gemm(L.down, ACTB, X, M, ACT_NONE, X, false, s, false, true, false, split_down, bn_down)) return rc;
    if (int rc = tap_stream(FVLA_TAP_LAYER0 + l)) return rc;
  }
  flop_scale_ = 1.0;
  // ---- final norm + pooling ----
  float* pooled = static_cast<float*>(ws_.bufs["pooled"].first);
  ++launches;
  prof_begin(s);
  if (int rc = pool_norm(DT_F32, X, final_norm_, d_pidx, d_lens, cfg.pool_mode, pooled, B, Tm, H,
                         cfg.rms_eps, s)) return rc;
  prof_end("llm.pool_norm", 0.0, static_cast<double>(B) * H * 8, s);
  if (int rc = tap(FVLA_TAP_POOLED, pooled, static_cast<size_t>(B) * H * 4, 0, s)) return rc;
  if (a.pooled != nullptr)
    FVLA_CUDA_CHECK(cudaMemcpyAsync(a.pooled, pooled, static_cast<size_t>(B) * H * 4,
                                    cudaMemcpyDeviceToDevice, s));
  // ---- action head ----
  if (a.states != nullptr) {
    FVLA_REQUIRE(a.actions != nullptr, "forward: actions buffer required with states");
    FVLA_REQUIRE(cfg.state_dim > 0, "forward: this engine was built without an action head");
    float* ts = static_cast<float*>(ws_.bufs["tap_state"].first);
    float* tf = static_cast<float*>(ws_.bufs["tap_fused"].first);
    ++launches;
    flops += 2.0 * B * (static_cast<double>(cfg.state_dim) * cfg.hidden_dim +
                        static_cast<double>(H + cfg.hidden_dim) * cfg.fusion_dim +
                        static_cast<double>(cfg.fusion_dim) * cfg.fusion_dim +
                        static_cast<double>(cfg.fusion_dim) * cfg.action_dim);
    prof_begin(s);
    float* tx = static_cast<float*>(ws_.bufs["head_x1"].first);
    if (int rc = action_head(cfg.dtype, head_, pooled, a.states, a.actions, ts, tx, tf, B, s)) return rc;
    prof_end("head.action_head", 0.0, 0.0, s);
    if (int rc = tap(FVLA_TAP_STATE_FEAT, ts, static_cast<size_t>(B) * cfg.hidden_dim * 4, 0, s)) return rc;
    if (int rc = tap(FVLA_TAP_FUSED, tf, static_cast<size_t>(B) * cfg.fusion_dim * 4, 0, s)) return rc;
  }
  return 0;
}

void Engine::clear_graphs() {
  for (auto& kv : graphs_)
    if (kv.second.exec != nullptr) cudaGraphExecDestroy(kv.second.exec);
  graphs_.clear();
}

// Inputs are copied into engine-owned staging buffers so that the captured launches never hold caller pointers;
// the first call with a new geometry runs eagerly (one-time attribute / static initialisation), the second is
// captured on an internal stream and instantiated, later calls are one cudaGraphLaunch into the caller's stream.
int Engine::forward_graph(const fvla_forward_args& a, int B, int Tm, bool any_image, cudaStream_t s) {
  const size_t img_bytes = a.images != nullptr
                               ? static_cast<size_t>(B) * a.img_c * a.img_h * a.img_w * dtype_size(a.img_dtype)
                               : 0;
  const size_t st_bytes = static_cast<size_t>(B) * cfg.state_dim * 4;
  const size_t act_bytes = static_cast<size_t>(B) * cfg.action_dim * 4;
  void *in_img = nullptr, *in_st = nullptr;
  if (img_bytes != 0)
    if (int rc = ensure("in_images", img_bytes, &in_img)) return rc;
  if (int rc = ensure("in_states", st_bytes, &in_st)) return rc;
  float* out_act = static_cast<float*>(ws_.bufs["actions"].first);

  fvla_forward_args g = a;
  g.images = in_img;
  g.states = static_cast<const float*>(in_st);
  g.actions = out_act;
  if (img_bytes != 0) FVLA_CUDA_CHECK(cudaMemcpyAsync(in_img, a.images, img_bytes, cudaMemcpyDeviceToDevice, s));
  FVLA_CUDA_CHECK(cudaMemcpyAsync(in_st, a.states, st_bytes, cudaMemcpyDeviceToDevice, s));

  // everything launch_all's control flow and kernel arguments depend on
  std::string key;
  auto add = [&key](const void* p, size_t n) { key.append(static_cast<const char*>(p), n); };
  const int ints[] = {B, Tm, any_image ? 1 : 0, a.img_c, a.img_h, a.img_w, a.img_dtype, a.img_nhwc, a.letterbox,
                      a.normalize, a.images != nullptr ? 1 : 0};
  add(ints, sizeof(ints));
  const float fl[] = {a.pad_value, a.img_scale, a.mean[0], a.mean[1], a.mean[2], a.inv_std[0], a.inv_std[1], a.inv_std[2]};
  add(fl, sizeof(fl));

  GraphEntry& ent = graphs_[key];
  int rc = 0;
  if (ent.exec != nullptr) {
    FVLA_CUDA_CHECK(cudaGraphLaunch(ent.exec, s));
    launches = ent.launches;
    flops = ent.flops;
  } else if (ent.state == 0 || ent.state == 3) {
    ent.state = ent.state == 0 ? 1 : 3;  // 1: warmed up eagerly, capture next time; 3: capture failed, stay eager
    rc = launch_all(g, B, Tm, any_image, s);
  } else {
    if (graph_stream_ == nullptr) FVLA_CUDA_CHECK(cudaStreamCreateWithFlags(&graph_stream_, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    cudaError_t ce = cudaStreamBeginCapture(graph_stream_, cudaStreamCaptureModeThreadLocal);
    if (ce == cudaSuccess) {
      rc = launch_all(g, B, Tm, any_image, graph_stream_);
      ce = cudaStreamEndCapture(graph_stream_, &graph);
    }
    if (ce == cudaSuccess && rc == 0 && graph != nullptr)
      ce = cudaGraphInstantiate(&ent.exec, graph, 0);
    if (graph != nullptr) cudaGraphDestroy(graph);
    if (ce != cudaSuccess || rc != 0 || ent.exec == nullptr) {
      cudaGetLastError();  // clear the sticky capture error, fall back to eager launches for this geometry
      ent.exec = nullptr;
      ent.state = 3;
      rc = launch_all(g, B, Tm, any_image, s);
    } else {
      ent.state = 2;
      ent.launches = launches;
      ent.flops = flops;
      FVLA_CUDA_CHECK(cudaGraphLaunch(ent.exec, s));
    }
  }
  if (rc != 0) return rc;
  FVLA_CUDA_CHECK(cudaMemcpyAsync(a.actions, out_act, act_bytes, cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // namespace fvla
