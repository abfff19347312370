// MultiViewUNet (Variant A) forward on B200: weight registry + packing, activation arena, and the
// straight-line launch sequence that replaces reference src/model/denoiser/mvunet.py:90-208 and
// src/model/denoiser/mvdream/attention.py:357-439 (+ diffusers ResnetBlock2D / Down/Upsample2D /
// Timesteps / TimestepEmbedding).
//
// Data layout in HBM: every activation is bf16 NHWC [image=(scene,view), h, w, C], i.e. a row-major
// [tokens, C] matrix, so conv3x3 / 1x1 / Linear are all `tokens x K` by `K x N` GEMMs (gemm_tc.cu) and the
// "b c h w -> b (h w) c" rearranges of the reference are no-ops.  Weights are bf16 [N, K] (K contiguous,
// tap-major for 3x3 filters).  Statistics, biases, time embeddings and the softmax are fp32.
#include <map>
#include <set>
#include <memory>
#include <vector>

#include "common.cuh"

namespace mvldm {

namespace {
thread_local std::string g_last_error;
}

void fail(const char* file, int line, const std::string& msg) {
  throw Error(std::string(file) + ":" + std::to_string(line) + ": " + msg);
}

__global__ void vec_add_kernel(float* dst, const float* a, const float* b, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = a[i] + (b ? b[i] : 0.f);
}

struct DevBuf {  // owning device allocation
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  explicit DevBuf(size_t n) { alloc(n); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  void alloc(size_t n) {
    release();
    if (n) MV_CUDA(cudaMalloc(&p, n));
    bytes = n;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  ~DevBuf() { release(); }
};

struct Act {  // bf16 NHWC activation view
  bf16* p = nullptr;
  int n = 0, h = 0, w = 0, c = 0;
  int64_t tokens() const { return (int64_t)n * h * w; }
};

struct Packed {  // bf16 [n, k] weight + fp32 bias
  bf16* w = nullptr;
  float* bias = nullptr;
  int n = 0, k = 0;
};

struct ResnetW {
  std::string key;
  int cin = 0, cout = 0, temb_off = 0;
  bool shortcut = false;
  bool temb = true;    // false: the VAE's ResnetBlock2D(temb_channels=None)
  float eps = 1e-5f;   // 1e-6 in the VAE
  const float *g1 = nullptr, *b1 = nullptr, *g2 = nullptr, *b2 = nullptr;
  Packed conv1, conv2;  // conv2 carries the 1x1 shortcut as extra K columns when present
};

struct MvW {  // SpatialTransformer3D block, or (t2d) a per-view diffusers Transformer2DModel of Variant B
  std::string key;
  int c = 0, d = 0, dpad = 0, heads = 0;
  bool t2d = false;
  const float *gn_g = nullptr, *gn_b = nullptr, *ln_g[3] = {}, *ln_b[3] = {};
  Packed proj_in, qkv1, out1, qkv2, out2, ff1, ff2, proj_out;
  // StandardTransformer (multi_view_attention.name == "standard"): `layers` pre-norm [joint attention, GELU MLP] pairs
  struct StdLayer {
    const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
    Packed qkv, out, fc1, fc2;
  };
  bool standard = false;
  int d_mlp = 0;
  std::vector<StdLayer> layers;
};

// AutoencoderKL mid-block attention (diffusers Attention, 1 head of C channels, GroupNorm 1e-6, residual)
struct VaeAttnW {
  std::string key;
  int c = 0;
  const float *gn_g = nullptr, *gn_b = nullptr;
  Packed q, k, v, out;
};
struct VaeHalf {  // encoder or decoder
  Packed conv_in, conv_out;
  std::vector<std::vector<ResnetW>> blocks;
  std::vector<Packed> resample;  // downsamplers.0.conv / upsamplers.0.conv per block (w == nullptr: none)
  ResnetW mid0, mid1;
  VaeAttnW attn;
  const float *norm_g = nullptr, *norm_b = nullptr;
};

struct Arena {
  char* base = nullptr;
  size_t off = 0, cap = 0, peak = 0;
  bool measuring = true;
  void* take(size_t bytes) {
    bytes = (bytes + 1023) & ~size_t(1023);  // 1 KB granularity keeps every buffer TMA/UMMA aligned
    void* r = measuring ? reinterpret_cast<void*>(size_t(0x100000) + off) : (void*)(base + off);
    off += bytes;
    if (off > peak) peak = off;
    if (!measuring) MV_CHECK(off <= cap, "activation arena overflow");
    return r;
  }
};

// One recorded forward for a batch shape: the launch list (every kernel with its arguments bound to the plan's buffers),
// the activation arena those point into and the CUDA graph of the whole list.
struct TapRec {
  Act a;
};
struct OpMeta {  // what a launch is, for the per-op timing report
  const char* cat;
  std::string what;
  double flops, bytes;
  double scores = 0, flops_padded = 0;  // attention only: softmax exponentials (one ex2 each) and tensor-pipe FLOPs on padded heads
};
struct Step {
  enum Kind { GEMM, GEMM_SIMT, ATTN, ATTN_SHARDED, GN, LN, UPSAMPLE, IM2COL, SINUSOID, SOFTMAX } kind = GEMM;
  mvldm_gemm_desc gemm{};  // GEMM / GEMM_SIMT
  // ATTN / ATTN_SHARDED
  const bf16* qkv = nullptr;
  bf16* out = nullptr;
  int batches = 0, seq = 0, heads = 0, d = 0, dpad = 0, seq_local = 0;
  bool simt = false;
  bf16* part[3] = {};     // ATTN_SHARDED: partial outputs (local slab, slabs before it, slabs after it) ...
  float* pstats[3] = {};  // ... and their softmax states, merged in that order
  // elementwise steps
  const void *x0 = nullptr, *x1 = nullptr;
  const float *gamma = nullptr, *beta = nullptr;
  void* dst = nullptr;
  float* scratch = nullptr;
  int c0 = 0, c1 = 0, n_img = 0, hw = 0, h = 0, w = 0, groups = 0, silu = 0, aux = 0;
  float eps = 0.f;
  // the time-embedding chain does not depend on the latents: it runs on a side lane (second stream / parallel graph branch)
  // next to im2col -> conv_in -> first GroupNorm, and the first launch that reads it joins the lanes again
  int lane = 0;
  bool join_side = false;
  OpMeta meta{};
};
struct Plan {
  bool side_joined = false;      // recording: the side lane has been joined back into the main one
  std::vector<int> scene_views;  // views per scene; images of a scene are contiguous
  int H = 0, W = 0;
  DevBuf arena_mem;
  size_t arena_bytes = 0;
  DevBuf in_latents, in_t, out_eps, splitk;
  std::vector<Step> steps;
  std::map<std::string, TapRec> taps;  // named intermediate activations (plans recorded with taps enabled keep them all alive)
  int launches = 0;  // kernels one execution launches
  uint64_t last_use = 0;
  cudaGraphExec_t graph = nullptr;
  ~Plan() {
    if (graph) cudaGraphExecDestroy(graph);
  }
};

}  // namespace mvldm

using namespace mvldm;

struct mvldm_handle_s {
  mvldm_config cfg{};
  int device = 0;
  int temb_dim = 0, temb_total = 0;
  bool finalized = false;

  // registry
  std::vector<std::string> names;
  std::map<std::string, std::vector<int64_t>> shapes;
  std::map<std::string, std::unique_ptr<DevBuf>> raw;  // fp32 copies of the state dict
  std::vector<std::unique_ptr<DevBuf>> packed_store;

  // packed layers
  Packed conv_in, conv_out, time1, time2, temb_all;
  std::vector<std::vector<ResnetW>> down_res, up_res;
  ResnetW mid_res;
  std::vector<Packed> down_conv, up_conv;
  std::vector<MvW> mv_enc, mv_dec;
  MvW mv_mid;
  // Variant B only
  std::vector<std::vector<MvW>> t2d_down;
  MvW t2d_mid;
  ResnetW mid_res1;
  std::set<std::string> unused;  // keys of modules the reference forward never calls (up-block attentions)
  const float *norm_out_g = nullptr, *norm_out_b = nullptr;
  int kpad_in = 0;
  // cfg.model == MVLDM_MODEL_VAE: AutoencoderKL (SD VAE) encoder / decoder on the same kernels
  VaeHalf vae_enc, vae_dec;
  const float* vae_premix = nullptr;  // post_quant_conv as [latent x latent weights | latent biases]
  int program = 0;                    // what plan_for records: 0 denoiser forward, 1 VAE decode, 2 VAE encode

  std::map<std::vector<int>, std::unique_ptr<Plan>> plans;
  int last_launches = 0;
  bool taps_enabled = false;
  Plan* last_plan = nullptr;

  // ---- profiling (mvldm_set_profiling): the launch list runs eagerly with a CUDA-event pair around every launch ----
  bool profiling = false;
  std::vector<cudaEvent_t> event_pool;
  std::string prof_json;
  Plan* prof_plan = nullptr;

  // ---- run state ----
  cudaStream_t stream = nullptr;
  cudaStream_t capture_stream = nullptr;  // graphs are recorded here: the caller's stream may be the legacy NULL stream
  cudaStream_t side_stream = nullptr;     // side lane of execute() (time-embedding chain)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  Arena arena;
  bool dry = true;       // measuring pass: arena offsets and split-K scratch only, nothing is recorded
  Plan* rec = nullptr;   // plan being recorded
  uint64_t use_clock = 0;
  // view-group sharding (mvldm_forward_sharded): local Q against all-gathered K/V in the joint attention
  bool sharded = false;
  int v_total = 0;
  void *kv_send = nullptr, *kv_recv = nullptr;
  size_t kv_recv_bytes = 0;
  mvldm_kv_exchange_fn exchange = nullptr;
  void* exchange_user = nullptr;
  int group_index = 0;  // which contiguous view group of the scene this rank holds
  size_t splitk_need = 0, splitk_bytes = 0;  // split-K fp32 scratch shared by all GEMMs of a forward
  void* splitk_ws = nullptr;

  // =========================== registry ===========================
  void reg(const std::string& k, std::vector<int64_t> shape) {
    names.push_back(k);
    shapes[k] = std::move(shape);
  }
  void reg_conv(const std::string& k, int cout, int cin, int ks) {
    reg(k + ".weight", {cout, cin, ks, ks});
    reg(k + ".bias", {cout});
  }
  void reg_lin(const std::string& k, int cout, int cin, bool bias = true) {
    reg(k + ".weight", {cout, cin});
    if (bias) reg(k + ".bias", {cout});
  }
  void reg_norm(const std::string& k, int c) {
    reg(k + ".weight", {c});
    reg(k + ".bias", {c});
  }
  ResnetW reg_resnet(const std::string& k, int cin, int cout) {
    reg_norm(k + ".norm1", cin);
    reg_conv(k + ".conv1", cout, cin, 3);
    reg_lin(k + ".time_emb_proj", cout, temb_dim);
    reg_norm(k + ".norm2", cout);
    reg_conv(k + ".conv2", cout, cout, 3);
    if (cin != cout) reg_conv(k + ".conv_shortcut", cout, cin, 1);
    ResnetW r;
    r.key = k;
    r.cin = cin;
    r.cout = cout;
    r.shortcut = cin != cout;
    r.temb_off = temb_total;
    temb_total += cout;
    return r;
  }
  // StandardTransformer state-dict keys: standard/transformer.py:76-84 -> transformer/transformer.py:24-47
  // (layers.J.0 = PreNorm(Attention: to_qkv no bias, to_out.0), layers.J.1 = PreNorm(FeedForward: net.0, GELU, net.3))
  MvW reg_standard(const std::string& k, int c) {
    MvW m;
    m.key = k;
    m.c = c;
    m.standard = true;
    m.heads = cfg.num_heads;
    MV_CHECK(c % cfg.num_heads == 0, "channels not divisible by num_heads");
    m.d = c / cfg.num_heads;
    m.dpad = (m.d + 63) / 64 * 64;
    MV_CHECK(m.dpad <= 192, "head dim > 192 unsupported");
    m.d_mlp = cfg.mv_d_mlp > 0 ? cfg.mv_d_mlp : c * cfg.mv_d_mlp_multiplier;
    MV_CHECK(m.d_mlp > 0 && m.d_mlp % 64 == 0, "standard transformer: d_mlp must be a positive multiple of 64");
    m.layers.resize(cfg.mv_num_layers);
    for (int j = 0; j < cfg.mv_num_layers; ++j) {
      const std::string lk = k + ".transformer.layers." + std::to_string(j);
      reg_norm(lk + ".0.norm", c);
      reg_lin(lk + ".0.fn.to_qkv", 3 * c, c, false);
      reg_lin(lk + ".0.fn.to_out.0", c, c);
      reg_norm(lk + ".1.norm", c);
      reg_lin(lk + ".1.fn.net.0", m.d_mlp, c);
      reg_lin(lk + ".1.fn.net.3", c, m.d_mlp);
    }
    return m;
  }
  MvW reg_mv(const std::string& k, int c) {
    if (cfg.mv_block == MVLDM_MV_STANDARD) return reg_standard(k, c);
    reg_norm(k + ".norm", c);
    reg_conv(k + ".proj_in", c, c, 1);
    const std::string tb = k + ".transformer_blocks.0";
    for (const char* a : {"attn1", "attn2"}) {
      reg_lin(tb + "." + a + ".to_q", c, c, false);
      reg_lin(tb + "." + a + ".to_k", c, c, false);
      reg_lin(tb + "." + a + ".to_v", c, c, false);
      reg_lin(tb + "." + a + ".to_out.0", c, c);
    }
    reg_lin(tb + ".ff.net.0.proj", 8 * c, c);
    reg_lin(tb + ".ff.net.2", c, 4 * c);
    for (const char* n : {"norm1", "norm2", "norm3"}) reg_norm(tb + "." + n, c);
    reg_conv(k + ".proj_out", c, c, 1);
    MvW m;
    m.key = k;
    m.c = c;
    m.heads = cfg.num_heads;
    MV_CHECK(c % cfg.num_heads == 0, "channels not divisible by num_heads");
    m.d = c / cfg.num_heads;
    m.dpad = (m.d + 63) / 64 * 64;
    MV_CHECK(m.dpad <= 192, "head dim > 192 unsupported");
    return m;
  }
  // diffusers Transformer2DModel (use_linear_projection) with one BasicTransformerBlock: SURVEY.md App. A.6
  MvW reg_t2d(const std::string& k, int c, int heads, bool used) {
    const size_t first = names.size();
    reg_norm(k + ".norm", c);
    reg_lin(k + ".proj_in", c, c);
    const std::string tb = k + ".transformer_blocks.0";
    reg_norm(tb + ".norm1", c);
    reg_lin(tb + ".attn1.to_q", c, c, false);
    reg_lin(tb + ".attn1.to_k", c, c, false);
    reg_lin(tb + ".attn1.to_v", c, c, false);
    reg_lin(tb + ".attn1.to_out.0", c, c);
    reg_norm(tb + ".norm2", c);
    reg_lin(tb + ".attn2.to_q", c, c, false);
    reg_lin(tb + ".attn2.to_k", c, cfg.cross_attention_dim, false);
    reg_lin(tb + ".attn2.to_v", c, cfg.cross_attention_dim, false);
    reg_lin(tb + ".attn2.to_out.0", c, c);
    reg_norm(tb + ".norm3", c);
    reg_lin(tb + ".ff.net.0.proj", 8 * c, c);
    reg_lin(tb + ".ff.net.2", c, 4 * c);
    reg_lin(k + ".proj_out", c, c);
    if (!used)
      for (size_t i = first; i < names.size(); ++i) unused.insert(names[i]);
    MvW m;
    m.key = k;
    m.c = c;
    m.t2d = true;
    m.heads = heads;
    MV_CHECK(heads > 0 && c % heads == 0, "t2d_heads must divide the level's channels");
    m.d = c / heads;
    m.dpad = (m.d + 63) / 64 * 64;  // d == dpad (SD-2.1: 64) runs the in-register row-sum attention kernel
    MV_CHECK(m.dpad <= 192, "head dim > 192 unsupported");
    return m;
  }

  void build_registry() {
    const int L = cfg.num_levels;
    const int* boc = cfg.block_out_channels;
    temb_dim = boc[0] * 4;
    reg_conv("unet.conv_in", boc[0], cfg.in_channels, 3);
    reg_lin("unet.time_embedding.linear_1", temb_dim, boc[0]);
    reg_lin("unet.time_embedding.linear_2", temb_dim, temb_dim);
    int cout = boc[0];
    down_res.resize(L);
    up_res.resize(L);
    t2d_down.resize(L);
    for (int l = 0; l < L; ++l) {
      const int cin = cout;
      cout = boc[l];
      for (int i = 0; i < cfg.layers_per_block; ++i)
        down_res[l].push_back(reg_resnet("unet.down_blocks." + std::to_string(l) + ".resnets." + std::to_string(i),
                                         i == 0 ? cin : cout, cout));
      if (cfg.variant == 1 && l != L - 1)
        for (int i = 0; i < cfg.layers_per_block; ++i)
          t2d_down[l].push_back(reg_t2d("unet.down_blocks." + std::to_string(l) + ".attentions." + std::to_string(i), cout,
                                        cfg.t2d_heads[l], true));
      if (l != L - 1) reg_conv("unet.down_blocks." + std::to_string(l) + ".downsamplers.0.conv", cout, cout, 3);
    }
    mid_res = reg_resnet("unet.mid_block.resnets.0", boc[L - 1], boc[L - 1]);
    if (cfg.variant == 1) {
      t2d_mid = reg_t2d("unet.mid_block.attentions.0", boc[L - 1], cfg.t2d_heads[L - 1], true);
      mid_res1 = reg_resnet("unet.mid_block.resnets.1", boc[L - 1], boc[L - 1]);
    }
    int out_c = boc[L - 1];
    for (int l = 0; l < L; ++l) {
      const int prev = out_c;
      out_c = boc[L - 1 - l];
      const int in_c = boc[L - 1 - std::min(l + 1, L - 1)];
      for (int i = 0; i < cfg.layers_per_block + 1; ++i) {
        const int skip = (i == cfg.layers_per_block) ? in_c : out_c;
        const int rin = (i == 0) ? prev : out_c;
        up_res[l].push_back(reg_resnet("unet.up_blocks." + std::to_string(l) + ".resnets." + std::to_string(i),
                                       rin + skip, out_c));
      }
      if (cfg.variant == 1 && l != 0)  // CrossAttnUpBlock2D attentions: in the state dict, never executed
        for (int i = 0; i < cfg.layers_per_block + 1; ++i)
          reg_t2d("unet.up_blocks." + std::to_string(l) + ".attentions." + std::to_string(i), out_c,
                  cfg.t2d_heads[L - 1 - l], false);
      if (l != L - 1) reg_conv("unet.up_blocks." + std::to_string(l) + ".upsamplers.0.conv", out_c, out_c, 3);
    }
    reg_norm("unet.conv_norm_out", boc[0]);
    reg_conv("unet.conv_out", cfg.out_channels, boc[0], 3);
    for (int l = 0; l < L; ++l) mv_enc.push_back(reg_mv("cross_attn_blocks_encoder." + std::to_string(l), boc[l]));
    mv_mid = reg_mv("cross_attn_blocks_mid.0", boc[L - 1]);
    for (int l = 0; l < L; ++l) mv_dec.push_back(reg_mv("cross_attn_blocks_decoder." + std::to_string(l), boc[L - 1 - l]));
  }

  // =========================== packing ===========================
  const float* rawf(const std::string& k) {
    auto it = raw.find(k);
    MV_CHECK(it != raw.end(), "weight not set: " + k);
    return reinterpret_cast<const float*>(it->second->p);
  }
  template <class T>
  T* store(size_t count) {
    packed_store.emplace_back(new DevBuf(count * sizeof(T)));
    MV_CUDA(cudaMemsetAsync(packed_store.back()->p, 0, count * sizeof(T), stream));
    return reinterpret_cast<T*>(packed_store.back()->p);
  }
  float* bias_sum(const std::string& a, const std::string& b, int n, int npad = 0) {
    float* dst = store<float>(std::max(n, npad));
    vec_add_kernel<<<ceil_div(n, 128), 128, 0, stream>>>(dst, rawf(a), b.empty() ? nullptr : rawf(b), n);
    MV_LAUNCHED();
    return dst;
  }
  const int* upload_map(const std::vector<int>& m) {
    int* d = store<int>(m.size());
    MV_CUDA(cudaMemcpyAsync(d, m.data(), m.size() * sizeof(int), cudaMemcpyHostToDevice, stream));
    MV_CUDA(cudaStreamSynchronize(stream));  // m may be a temporary
    return d;
  }
  Packed pack_conv(const std::string& k, int cout, int cin, int ks, int extra_k = 0, int npad = 0) {
    Packed p;
    p.n = std::max(cout, npad);
    p.k = ks * ks * cin + extra_k;
    p.w = store<bf16>((size_t)p.n * p.k);
    pack_conv3x3(stream, rawf(k + ".weight"), cout, cin, ks, p.w, p.k, 0);
    p.bias = bias_sum(k + ".bias", "", cout, p.n);
    return p;
  }
  Packed pack_linear(const std::string& k, int n, int kk, bool bias) {
    Packed p;
    p.n = n;
    p.k = kk;
    p.w = store<bf16>((size_t)n * kk);
    pack_rows(stream, rawf(k + ".weight"), n, kk, kk, p.w, kk, 0, nullptr);
    if (bias) p.bias = bias_sum(k + ".bias", "", n);
    return p;
  }
  void pack_resnet(ResnetW& r) {
    r.g1 = rawf(r.key + ".norm1.weight");
    r.b1 = rawf(r.key + ".norm1.bias");
    r.g2 = rawf(r.key + ".norm2.weight");
    r.b2 = rawf(r.key + ".norm2.bias");
    r.conv1 = pack_conv(r.key + ".conv1", r.cout, r.cin, 3);
    if (r.shortcut) {
      r.conv2 = pack_conv(r.key + ".conv2", r.cout, r.cout, 3, r.cin);
      pack_rows(stream, rawf(r.key + ".conv_shortcut.weight"), r.cout, r.cin, r.cin, r.conv2.w, r.conv2.k, 9 * r.cout,
                nullptr);
      r.conv2.bias = bias_sum(r.key + ".conv2.bias", r.key + ".conv_shortcut.bias", r.cout);
    } else {
      r.conv2 = pack_conv(r.key + ".conv2", r.cout, r.cout, 3);
    }
    // rows [temb_off, temb_off + cout) of the batched time_emb_proj
    pack_rows(stream, rawf(r.key + ".time_emb_proj.weight"), r.cout, temb_dim, temb_dim,
              temb_all.w + (size_t)r.temb_off * temb_dim, temb_dim, 0, nullptr);
    MV_CUDA(cudaMemcpyAsync(temb_all.bias + r.temb_off, rawf(r.key + ".time_emb_proj.bias"), r.cout * sizeof(float),
                            cudaMemcpyDeviceToDevice, stream));
  }
  // `fused`: the source is ONE [3C, C] matrix (transformer/attention.py:47, chunk(3) = q | k | v rows) instead of the
  // three to_q / to_k / to_v matrices of mvdream/attention.py:168-170
  Packed pack_qkv(const std::string& a, const MvW& m, bool fused = false) {
    const int H = m.heads;
    Packed p;
    p.n = 3 * H * m.dpad;
    p.k = m.c;
    p.w = store<bf16>((size_t)p.n * p.k);
    const char* which[3] = {".to_q", ".to_k", ".to_v"};
    for (int t = 0; t < 3; ++t) {
      std::vector<int> map(m.c);
      for (int r = 0; r < m.c; ++r) map[r] = (t * H + r / m.d) * m.dpad + r % m.d;
      const float* src = fused ? rawf(a + ".to_qkv.weight") + (size_t)t * m.c * m.c : rawf(a + which[t] + ".weight");
      pack_rows(stream, src, m.c, m.c, m.c, p.w, p.k, 0, upload_map(map));
    }
    // q/k/v have no bias in the reference; the packed GEMM's bias plants 1.0 in the first pad column of every V head,
    // which makes P.V deliver the softmax row sum in column d of the attention accumulator (attn_tc.cu)
    // (d == dpad, Variant B's 64-wide heads: no pad column; attention_tc then sums the row in registers)
    std::vector<float> hb(p.n, 0.f);
    if (m.dpad > m.d)
      for (int h = 0; h < H; ++h) hb[(2 * H + h) * m.dpad + m.d] = 1.f;
    p.bias = store<float>(p.n);
    MV_CUDA(cudaMemcpyAsync(p.bias, hb.data(), hb.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
    MV_CUDA(cudaStreamSynchronize(stream));
    return p;
  }
  Packed pack_attn_out(const std::string& a, const MvW& m) {
    const int H = m.heads;
    Packed p;
    p.n = m.c;
    p.k = H * m.dpad;
    p.w = store<bf16>((size_t)p.n * p.k);
    for (int h = 0; h < H; ++h)
      pack_rows(stream, rawf(a + ".to_out.0.weight") + h * m.d, m.c, m.d, m.c, p.w, p.k, h * m.dpad, nullptr);
    p.bias = bias_sum(a + ".to_out.0.bias", "", m.c);
    return p;
  }
  void pack_standard(MvW& m) {
    for (size_t j = 0; j < m.layers.size(); ++j) {
      const std::string lk = m.key + ".transformer.layers." + std::to_string(j);
      MvW::StdLayer& y = m.layers[j];
      y.ln1_g = rawf(lk + ".0.norm.weight");
      y.ln1_b = rawf(lk + ".0.norm.bias");
      y.ln2_g = rawf(lk + ".1.norm.weight");
      y.ln2_b = rawf(lk + ".1.norm.bias");
      y.qkv = pack_qkv(lk + ".0.fn", m, true);
      y.out = pack_attn_out(lk + ".0.fn", m);
      y.fc1 = pack_linear(lk + ".1.fn.net.0", m.d_mlp, m.c, true);
      y.fc2 = pack_linear(lk + ".1.fn.net.3", m.c, m.d_mlp, true);
    }
  }
  void pack_mv(MvW& m) {
    if (m.standard) return pack_standard(m);
    const std::string tb = m.key + ".transformer_blocks.0";
    m.gn_g = rawf(m.key + ".norm.weight");
    m.gn_b = rawf(m.key + ".norm.bias");
    const char* ln[3] = {".norm1", ".norm2", ".norm3"};
    for (int i = 0; i < 3; ++i) {
      m.ln_g[i] = rawf(tb + ln[i] + ".weight");
      m.ln_b[i] = rawf(tb + ln[i] + ".bias");
    }
    m.qkv1 = pack_qkv(tb + ".attn1", m);
    m.out1 = pack_attn_out(tb + ".attn1", m);
    if (m.t2d) {
      m.proj_in = pack_linear(m.key + ".proj_in", m.c, m.c, true);
      m.proj_out = pack_linear(m.key + ".proj_out", m.c, m.c, true);
      // attn2 cross-attends to ONE all-zero token (mvunet.py:125-128): K = V = W.0 = 0 exactly, softmax over one key
      // is 1, so attn2(x) == to_out.0.bias for every token whatever its to_q/to_k/to_v/norm2 hold.  The bias rides
      // on the attn1 output projection.
      m.out1.bias = bias_sum(tb + ".attn1.to_out.0.bias", tb + ".attn2.to_out.0.bias", m.c);
    } else {
      m.proj_in = pack_conv(m.key + ".proj_in", m.c, m.c, 1);
      m.proj_out = pack_conv(m.key + ".proj_out", m.c, m.c, 1);
      m.qkv2 = pack_qkv(tb + ".attn2", m);
      m.out2 = pack_attn_out(tb + ".attn2", m);
    }
    // GEGLU: interleave value / gate rows in blocks of 8 so one 16-column accumulator chunk holds both
    const int c4 = 4 * m.c;
    m.ff1.n = 2 * c4;
    m.ff1.k = m.c;
    m.ff1.w = store<bf16>((size_t)m.ff1.n * m.ff1.k);
    std::vector<int> map(2 * c4);
    for (int ch = 0; ch < c4; ++ch) {
      map[ch] = (ch / 8) * 16 + ch % 8;
      map[c4 + ch] = (ch / 8) * 16 + 8 + ch % 8;
    }
    pack_rows(stream, rawf(tb + ".ff.net.0.proj.weight"), 2 * c4, m.c, m.c, m.ff1.w, m.c, 0, upload_map(map));
    {
      std::vector<float> hb(2 * c4), pb(2 * c4);
      MV_CUDA(cudaMemcpyAsync(hb.data(), rawf(tb + ".ff.net.0.proj.bias"), hb.size() * sizeof(float),
                              cudaMemcpyDeviceToHost, stream));
      MV_CUDA(cudaStreamSynchronize(stream));
      for (int i = 0; i < 2 * c4; ++i) pb[map[i]] = hb[i];
      m.ff1.bias = store<float>(2 * c4);
      MV_CUDA(cudaMemcpyAsync(m.ff1.bias, pb.data(), pb.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
      MV_CUDA(cudaStreamSynchronize(stream));
    }
    // ff.net.2 and proj_out are both linear with nothing in between:  proj_out(t3 + W2 f + b2) = (Wp W2) f + Wp t3 + (Wp b2 + bp).
    // ONE GEMM over the K-segments [f (4C) | t3 (C)] with the pre-multiplied matrix - same FLOPs as the two it replaces, one
    // launch and one bf16 round trip of the [tokens, C] intermediate less per block
    {
      const float* wp = rawf(m.key + ".proj_out.weight");   // [C, C] (1x1 conv or Linear: same layout)
      const float* w2 = rawf(tb + ".ff.net.2.weight");      // [C, 4C]
      float* prod = store<float>((size_t)m.c * c4);
      matmul_f32(stream, wp, w2, prod, m.c, c4, m.c);
      m.ff2.n = m.c;
      m.ff2.k = c4 + m.c;
      m.ff2.w = store<bf16>((size_t)m.ff2.n * m.ff2.k);
      pack_rows(stream, prod, m.c, c4, c4, m.ff2.w, m.ff2.k, 0, nullptr);
      pack_rows(stream, wp, m.c, m.c, m.c, m.ff2.w, m.ff2.k, c4, nullptr);
      m.ff2.bias = store<float>(m.c);
      matvec_bias(stream, wp, rawf(tb + ".ff.net.2.bias"), rawf(m.key + ".proj_out.bias"), m.ff2.bias, m.c, m.c);
    }
  }

  void finalize(cudaStream_t s) {
    stream = s;
    for (auto& n : names) MV_CHECK(unused.count(n) || raw.count(n), "mvldm_finalize_weights: missing weight " + n);
    MV_CUDA(cudaStreamSynchronize(s));
    packed_store.clear();
    const int L = cfg.num_levels;
    const int* boc = cfg.block_out_channels;
    kpad_in = (9 * cfg.in_channels + 63) / 64 * 64;
    {  // conv_in on the explicit im2col operand, K padded to a multiple of 64
      conv_in.n = boc[0];
      conv_in.k = kpad_in;
      conv_in.w = store<bf16>((size_t)conv_in.n * conv_in.k);
      pack_conv3x3(stream, rawf("unet.conv_in.weight"), boc[0], cfg.in_channels, 3, conv_in.w, conv_in.k, 0);
      conv_in.bias = bias_sum("unet.conv_in.bias", "", boc[0]);
    }
    conv_out = pack_conv("unet.conv_out", cfg.out_channels, boc[0], 3, 0, 32);
    time1 = pack_linear("unet.time_embedding.linear_1", temb_dim, boc[0], true);
    time2 = pack_linear("unet.time_embedding.linear_2", temb_dim, temb_dim, true);
    temb_all.n = temb_total;
    temb_all.k = temb_dim;
    temb_all.w = store<bf16>((size_t)temb_total * temb_dim);
    temb_all.bias = store<float>(temb_total);
    for (auto& lv : down_res)
      for (auto& r : lv) pack_resnet(r);
    pack_resnet(mid_res);
    if (cfg.variant == 1) {
      pack_resnet(mid_res1);
      for (auto& lv : t2d_down)
        for (auto& m : lv) pack_mv(m);
      pack_mv(t2d_mid);
    }
    for (auto& lv : up_res)
      for (auto& r : lv) pack_resnet(r);
    down_conv.clear();
    up_conv.clear();
    for (int l = 0; l < L - 1; ++l) {
      down_conv.push_back(pack_conv("unet.down_blocks." + std::to_string(l) + ".downsamplers.0.conv", boc[l], boc[l], 3));
      up_conv.push_back(pack_conv("unet.up_blocks." + std::to_string(l) + ".upsamplers.0.conv", boc[L - 1 - l],
                                  boc[L - 1 - l], 3));
    }
    for (auto& m : mv_enc) pack_mv(m);
    pack_mv(mv_mid);
    for (auto& m : mv_dec) pack_mv(m);
    norm_out_g = rawf("unet.conv_norm_out.weight");
    norm_out_b = rawf("unet.conv_norm_out.bias");
    MV_CUDA(cudaStreamSynchronize(s));
    // keep only the norm affine params of the fp32 copies
    for (auto it = raw.begin(); it != raw.end();) {
      const std::string& k = it->first;
      const bool keep = k.find("norm") != std::string::npos;
      it = keep ? std::next(it) : raw.erase(it);
    }
    plans.clear();
    prof_plan = last_plan = nullptr;
    finalized = true;
  }

  // =========================== forward ===========================
  Act new_act(int n, int h, int w, int c) {
    Act a;
    a.n = n; a.h = h; a.w = w; a.c = c;
    a.p = reinterpret_cast<bf16*>(arena.take((size_t)a.tokens() * c * sizeof(bf16)));
    return a;
  }
  float* new_f32(size_t count) { return reinterpret_cast<float*>(arena.take(count * sizeof(float))); }
  void tap(const std::string& name, const Act& a) {
    if (taps_enabled && !dry) rec->taps[name].a = a;
  }

  static mvldm_aseg seg_conv3x3(const Act& a, int stride = 1) {
    mvldm_aseg s{};
    s.ptr = a.p; s.c = a.c; s.ctot = a.c; s.sh = a.h; s.sw = a.w; s.stride = stride; s.ntaps = 9;
    for (int t = 0; t < 9; ++t) {
      s.dh[t] = (int8_t)(t / 3 - 1);
      s.dw[t] = (int8_t)(t % 3 - 1);
      s.coff[t] = 0;
    }
    return s;
  }
  static mvldm_aseg seg_1x1(const Act& a) {
    mvldm_aseg s{};
    s.ptr = a.p; s.c = a.c; s.ctot = a.c; s.sh = a.h; s.sw = a.w; s.stride = 1; s.ntaps = 1;
    return s;
  }
  // ---- recording: every op of the forward lands in the plan's launch list ----
  void push_step(const Step& st) { rec->steps.push_back(st); }
  void run_gemm(mvldm_gemm_desc& d, double algo_flops = -1.0) {
    const double M = (double)d.n_img * d.oh * d.ow;
    const char* cat = d.nseg > 0 && d.seg[0].ntaps == 9 ? "gemm_conv3x3" : "gemm_linear";
    if (dry) {
      if (cfg.impl != MVLDM_IMPL_SIMT) splitk_need = std::max(splitk_need, gemm_tc_workspace_bytes(d));
      return;
    }
    Step st;
    st.kind = cfg.impl == MVLDM_IMPL_SIMT ? Step::GEMM_SIMT : Step::GEMM;
    st.gemm = d;
    if (d.rowvec && !rec->side_joined) {  // first reader of the time embedding
      st.join_side = true;
      rec->side_joined = true;
    }
    // algorithmic bytes: weights once + every input segment once + the output (+ the residual), no re-reads
    double bytes = 2.0 * (double)d.n * d.k + (d.mode == 2 ? 4.0 * M * d.n_valid : 2.0 * M * d.n * (d.mode == 1 ? 0.5 : 1.0));
    for (int i = 0; i < d.nseg; ++i) bytes += 2.0 * (double)d.n_img * d.seg[i].sh * d.seg[i].sw * d.seg[i].c;
    if (d.residual) bytes += 2.0 * M * d.n;
    st.meta = OpMeta{cat, "M" + std::to_string((long long)M) + " N" + std::to_string(d.n) + " K" + std::to_string(d.k),
                     algo_flops >= 0 ? algo_flops : 2.0 * M * d.n * d.k, bytes};
    push_step(st);
  }
  // out = A-segments x W^T (+bias +rowvec +residual), bf16 NHWC
  void gemm(std::initializer_list<mvldm_aseg> segs, const Packed& w, const Act& out, const float* rowvec = nullptr,
            int rowvec_ld = 0, const Act* residual = nullptr, int mode = 0, double algo_flops = -1.0) {
    mvldm_gemm_desc d{};
    for (const auto& s : segs) d.seg[d.nseg++] = s;
    d.n_img = out.n; d.oh = out.h; d.ow = out.w;
    d.w = w.w; d.n = w.n; d.k = w.k;
    d.bias = w.bias;
    d.rowvec = rowvec; d.rowvec_ld = rowvec_ld;
    if (residual) { d.residual = residual->p; d.res_ld = residual->c; }
    d.mode = mode; d.out = out.p; d.ldo = out.c; d.n_valid = w.n;
    run_gemm(d, algo_flops);
  }
  void gn(const Act& x0, const Act* x1, const float* g, const float* b, float eps, bool silu, const Act& out) {
    float* scratch = new_f32(groupnorm_scratch_floats(x0.n, cfg.norm_groups));
    if (dry) return;
    Step st;
    st.kind = Step::GN;
    st.x0 = x0.p; st.c0 = x0.c; st.x1 = x1 ? x1->p : nullptr; st.c1 = x1 ? x1->c : 0;
    st.n_img = x0.n; st.hw = x0.h * x0.w; st.groups = cfg.norm_groups; st.eps = eps; st.gamma = g; st.beta = b;
    st.silu = silu ? 1 : 0; st.dst = out.p; st.scratch = scratch;
    st.meta = OpMeta{"groupnorm", "tokens" + std::to_string(x0.tokens()) + " C" + std::to_string(out.c), 0.0,
                     4.0 * (double)out.tokens() * out.c};
    push_step(st);
  }
  void ln(const Act& x, const float* g, const float* b, const Act& out) {
    if (dry) return;
    Step st;
    st.kind = Step::LN;
    st.x0 = x.p; st.n_img = (int)x.tokens(); st.c0 = x.c; st.eps = 1e-5f; st.gamma = g; st.beta = b; st.dst = out.p;
    st.meta = OpMeta{"layernorm", "tokens" + std::to_string(x.tokens()) + " C" + std::to_string(x.c), 0.0,
                     4.0 * (double)x.tokens() * x.c};
    push_step(st);
  }
  void attn(const Act& qkv, const Act& out, int batches, int seq, const MvW& m) {
    if (dry) return;
    Step st;
    st.kind = Step::ATTN;
    st.qkv = qkv.p; st.out = out.p; st.batches = batches; st.seq = seq; st.heads = m.heads; st.d = m.d; st.dpad = m.dpad;
    st.simt = cfg.impl != MVLDM_IMPL_TC;
    // algorithmic FLOPs: QK^T and PV over the un-padded head dim, 4 * seq^2 * C per batch
    st.meta = OpMeta{seq > out.h * out.w ? "attention_joint" : "attention_per_view",
                     "batches" + std::to_string(batches) + " seq" + std::to_string(seq) + " d" + std::to_string(m.d),
                     4.0 * batches * (double)seq * seq * m.c, 2.0 * (double)out.tokens() * (4.0 * m.heads * m.dpad)};
    st.meta.scores = (double)batches * m.heads * (double)seq * seq;
    st.meta.flops_padded = 4.0 * batches * (double)seq * seq * m.heads * m.dpad;
    rec->steps.push_back(st);
  }

  // This rank's views supply the queries; K and V of every view of the scene are all-gathered by the host callback.
  void joint_attention_sharded(const Act& qkv, const Act& out, int seq_local, const MvW& m) {
    const int world = v_total / (int)(qkv.n);
    // up to three partial softmaxes (own slab while the exchange is in flight, the slabs before, the slabs after)
    Act parts[3];
    float* stats[3] = {};
    if (world > 1)
      for (int i = 0; i < 3; ++i) {
        parts[i] = new_act(out.n, out.h, out.w, out.c);
        stats[i] = new_f32((size_t)seq_local * m.heads * 2);
      }
    if (dry) return;
    Step st;
    st.kind = Step::ATTN_SHARDED;
    st.qkv = qkv.p; st.out = out.p; st.batches = 1; st.seq = seq_local * world; st.seq_local = seq_local;
    st.heads = m.heads; st.d = m.d; st.dpad = m.dpad;
    for (int i = 0; i < 3; ++i) {
      st.part[i] = parts[i].p;
      st.pstats[i] = stats[i];
    }
    st.meta = OpMeta{"attention_joint_sharded", "seq_q" + std::to_string(seq_local) + " seq_kv" + std::to_string(seq_local * world),
                     4.0 * (double)seq_local * seq_local * world * m.c, 0.0};
    st.meta.scores = (double)m.heads * seq_local * (double)seq_local * world;
    st.meta.flops_padded = 4.0 * (double)seq_local * seq_local * world * m.heads * m.dpad;
    rec->steps.push_back(st);
  }

  Act resnet(const ResnetW& r, const Act& x0, const Act* x1, const float* temb) {
    const int n = x0.n, h = x0.h, w = x0.w;
    MV_CHECK(x0.c + (x1 ? x1->c : 0) == r.cin, "resnet channel mismatch");
    Act out = new_act(n, h, w, r.cout);
    const size_t mark = arena.off;
    Act a = new_act(n, h, w, r.cin);
    gn(x0, x1, r.g1, r.b1, r.eps, true, a);
    Act hdn = new_act(n, h, w, r.cout);
    if (r.temb) gemm({seg_conv3x3(a)}, r.conv1, hdn, temb + r.temb_off, temb_total);
    else gemm({seg_conv3x3(a)}, r.conv1, hdn);
    Act a2 = new_act(n, h, w, r.cout);
    gn(hdn, nullptr, r.g2, r.b2, r.eps, true, a2);
    if (r.shortcut) {
      if (x1) gemm({seg_conv3x3(a2), seg_1x1(x0), seg_1x1(*x1)}, r.conv2, out);
      else gemm({seg_conv3x3(a2), seg_1x1(x0)}, r.conv2, out);
    } else {
      MV_CHECK(!x1, "identity-shortcut resnet cannot take a concat input");
      gemm({seg_conv3x3(a2)}, r.conv2, out, nullptr, 0, &x0);
    }
    if (!taps_enabled) arena.off = mark;
    return out;
  }

  // joint attention per scene; runs of scenes with the same view count share one launch (uniform batches: one launch)
  void joint_attention(const Act& qkv, const Act& o, const MvW& m) {
    const int hw = qkv.h * qkv.w;
    size_t i = 0;
    int img0 = 0;
    while (i < scene_views.size()) {
      size_t j = i;
      while (j < scene_views.size() && scene_views[j] == scene_views[i]) ++j;
      const int V = scene_views[i], cnt = (int)(j - i);
      Act q = qkv, oo = o;
      q.p = qkv.p + (size_t)img0 * hw * qkv.c;
      oo.p = o.p + (size_t)img0 * hw * o.c;
      q.n = oo.n = cnt * V;
      attn(q, oo, cnt, V * hw, m);
      img0 += cnt * V;
      i = j;
    }
  }

  // StandardTransformer.forward (standard/transformer.py:96-136 with downscale 1, pos_enc off) around
  // Transformer.forward (transformer/transformer.py:68-72): tokens "b v c h w -> b (v h w) c" are this library's NHWC rows
  Act standard_block(const MvW& m, const Act& x) {
    const int n = x.n, h = x.h, w = x.w, C = m.c, H = m.heads;
    const int hw = h * w;
    Act cur = x;
    Act out = new_act(n, h, w, C);
    const size_t mark = arena.off;
    Act nrm = new_act(n, h, w, C);
    Act qkv = new_act(n, h, w, 3 * H * m.dpad);
    Act o = new_act(n, h, w, H * m.dpad);
    Act f = new_act(n, h, w, m.d_mlp);
    for (size_t j = 0; j < m.layers.size(); ++j) {
      const MvW::StdLayer& y = m.layers[j];
      const bool last = j + 1 == m.layers.size();
      ln(cur, y.ln1_g, y.ln1_b, nrm);
      gemm({seg_1x1(nrm)}, y.qkv, qkv, nullptr, 0, nullptr, 0, 6.0 * (double)x.tokens() * C * C);
      if (sharded) joint_attention_sharded(qkv, o, scene_views[0] * hw, m);
      else joint_attention(qkv, o, m);
      Act t1 = new_act(n, h, w, C);
      gemm({seg_1x1(o)}, y.out, t1, nullptr, 0, &cur, 0, 2.0 * (double)x.tokens() * C * C);
      if (j == 0) tap(m.key + ".attn", t1);
      ln(t1, y.ln2_g, y.ln2_b, nrm);
      gemm({seg_1x1(nrm)}, y.fc1, f, nullptr, 0, nullptr, 5);
      Act t2 = last ? out : new_act(n, h, w, C);
      gemm({seg_1x1(f)}, y.fc2, t2, nullptr, 0, &t1);
      cur = t2;
    }
    if (!taps_enabled) arena.off = mark;
    return out;
  }

  Act mv_block(const MvW& m, const Act& x) {
    if (m.standard) return standard_block(m, x);
    const int n = x.n, h = x.h, w = x.w, C = m.c, H = m.heads;
    const int hw = h * w;
    Act out = new_act(n, h, w, C);
    const size_t mark = arena.off;
    Act g = new_act(n, h, w, C);
    gn(x, nullptr, m.gn_g, m.gn_b, 1e-6f, false, g);
    Act t = new_act(n, h, w, C);
    gemm({seg_1x1(g)}, m.proj_in, t);
    Act nrm = new_act(n, h, w, C);
    Act qkv = new_act(n, h, w, 3 * H * m.dpad);
    Act o = new_act(n, h, w, H * m.dpad);
    // joint attention over all V*h*w tokens of a scene ("(b f) l c -> b (f l) c")
    ln(t, m.ln_g[0], m.ln_b[0], nrm);
    gemm({seg_1x1(nrm)}, m.qkv1, qkv, nullptr, 0, nullptr, 0, 6.0 * (double)t.tokens() * C * C);
    if (sharded) joint_attention_sharded(qkv, o, scene_views[0] * hw, m);
    else joint_attention(qkv, o, m);
    Act t2 = new_act(n, h, w, C);
    gemm({seg_1x1(o)}, m.out1, t2, nullptr, 0, &t, 0, 2.0 * (double)t.tokens() * C * C);
    tap(m.key + ".attn1", t2);
    // per-view attention
    ln(t2, m.ln_g[1], m.ln_b[1], nrm);
    gemm({seg_1x1(nrm)}, m.qkv2, qkv, nullptr, 0, nullptr, 0, 6.0 * (double)t.tokens() * C * C);
    attn(qkv, o, n, hw, m);
    Act t3 = new_act(n, h, w, C);
    gemm({seg_1x1(o)}, m.out2, t3, nullptr, 0, &t2, 0, 2.0 * (double)t.tokens() * C * C);
    tap(m.key + ".attn2", t3);
    // GEGLU feed-forward
    ln(t3, m.ln_g[2], m.ln_b[2], nrm);
    Act f = new_act(n, h, w, 4 * C);
    gemm({seg_1x1(nrm)}, m.ff1, f, nullptr, 0, nullptr, 1);
    gemm({seg_1x1(f), seg_1x1(t3)}, m.ff2, out, nullptr, 0, &x);  // ff.net.2 + residual + proj_out + residual (see pack_mv)
    if (!taps_enabled) arena.off = mark;
    return out;
  }

  // Variant B: per-view diffusers Transformer2DModel (GN 1e-6 -> Linear -> [LN, self-attn, +] -> [+ attn2 bias]
  // -> [LN, GEGLU FF, +] -> Linear -> + x); the attention never crosses views (mvunet.py:118-131).
  Act t2d_block(const MvW& m, const Act& x) {
    const int n = x.n, h = x.h, w = x.w, C = m.c, H = m.heads;
    Act out = new_act(n, h, w, C);
    const size_t mark = arena.off;
    Act g = new_act(n, h, w, C);
    gn(x, nullptr, m.gn_g, m.gn_b, 1e-6f, false, g);
    Act t = new_act(n, h, w, C);
    gemm({seg_1x1(g)}, m.proj_in, t);
    Act nrm = new_act(n, h, w, C);
    Act qkv = new_act(n, h, w, 3 * H * m.dpad);
    Act o = new_act(n, h, w, H * m.dpad);
    ln(t, m.ln_g[0], m.ln_b[0], nrm);
    gemm({seg_1x1(nrm)}, m.qkv1, qkv, nullptr, 0, nullptr, 0, 6.0 * (double)t.tokens() * C * C);
    attn(qkv, o, n, h * w, m);
    Act t3 = new_act(n, h, w, C);  // attn1 residual + the constant attn2 output (see pack_mv)
    gemm({seg_1x1(o)}, m.out1, t3, nullptr, 0, &t, 0, 2.0 * (double)t.tokens() * C * C);
    ln(t3, m.ln_g[2], m.ln_b[2], nrm);
    Act f = new_act(n, h, w, 4 * C);
    gemm({seg_1x1(nrm)}, m.ff1, f, nullptr, 0, nullptr, 1);
    gemm({seg_1x1(f), seg_1x1(t3)}, m.ff2, out, nullptr, 0, &x);  // ff.net.2 + residual + proj_out + residual (see pack_mv)
    if (!taps_enabled) arena.off = mark;
    return out;
  }


  // =========================== AutoencoderKL (cfg.model == MVLDM_MODEL_VAE) ===========================
  // diffusers AutoencoderKL as the reference constructs it (src/model/autoencoder/__init__.py:40-43) and calls it
  // (diffusion_wrapper.py:278-298): Encoder / Decoder of ResnetBlock2D(temb=None, eps 1e-6) stacks, one single-head attention
  // in each mid block, quant_conv / post_quant_conv.  State-dict keys are diffusers' (encoder.*, decoder.*, quant_conv.*,
  // post_quant_conv.*).  Same kernels as the denoiser: implicit-GEMM conv3x3, GroupNorm(+SiLU), nearest 2x upsample.
  ResnetW reg_vae_resnet(const std::string& k, int cin, int cout) {
    reg_norm(k + ".norm1", cin);
    reg_conv(k + ".conv1", cout, cin, 3);
    reg_norm(k + ".norm2", cout);
    reg_conv(k + ".conv2", cout, cout, 3);
    if (cin != cout) reg_conv(k + ".conv_shortcut", cout, cin, 1);
    ResnetW r;
    r.key = k; r.cin = cin; r.cout = cout; r.shortcut = cin != cout; r.temb = false; r.eps = 1e-6f;
    return r;
  }
  VaeAttnW reg_vae_attn(const std::string& k, int c) {
    reg_norm(k + ".group_norm", c);
    for (const char* n : {".to_q", ".to_k", ".to_v", ".to_out.0"}) reg_lin(k + n, c, c);
    VaeAttnW a;
    a.key = k; a.c = c;
    return a;
  }
  void build_vae_registry() {
    const int L = cfg.num_levels;
    const int* boc = cfg.block_out_channels;
    const int lc = cfg.latent_channels;
    // ---- encoder: conv_in, L DownEncoderBlock2D (layers_per_block resnets, stride-2 conv except the last), mid, head
    reg_conv("encoder.conv_in", boc[0], cfg.in_channels, 3);
    vae_enc.blocks.resize(L);
    int c = boc[0];
    for (int l = 0; l < L; ++l) {
      for (int i = 0; i < cfg.layers_per_block; ++i) {
        vae_enc.blocks[l].push_back(reg_vae_resnet("encoder.down_blocks." + std::to_string(l) + ".resnets." + std::to_string(i),
                                                   i == 0 ? c : boc[l], boc[l]));
      }
      c = boc[l];
      if (l != L - 1) reg_conv("encoder.down_blocks." + std::to_string(l) + ".downsamplers.0.conv", c, c, 3);
    }
    vae_enc.mid0 = reg_vae_resnet("encoder.mid_block.resnets.0", c, c);
    vae_enc.attn = reg_vae_attn("encoder.mid_block.attentions.0", c);
    vae_enc.mid1 = reg_vae_resnet("encoder.mid_block.resnets.1", c, c);
    reg_norm("encoder.conv_norm_out", c);
    reg_conv("encoder.conv_out", 2 * lc, c, 3);
    reg_conv("quant_conv", 2 * lc, 2 * lc, 1);
    reg_conv("post_quant_conv", lc, lc, 1);
    // ---- decoder: conv_in, mid, L UpDecoderBlock2D (layers_per_block + 1 resnets, nearest-2x + conv except the last), head
    reg_conv("decoder.conv_in", boc[L - 1], lc, 3);
    c = boc[L - 1];
    vae_dec.mid0 = reg_vae_resnet("decoder.mid_block.resnets.0", c, c);
    vae_dec.attn = reg_vae_attn("decoder.mid_block.attentions.0", c);
    vae_dec.mid1 = reg_vae_resnet("decoder.mid_block.resnets.1", c, c);
    vae_dec.blocks.resize(L);
    for (int l = 0; l < L; ++l) {
      const int co = boc[L - 1 - l];
      for (int i = 0; i < cfg.layers_per_block + 1; ++i)
        vae_dec.blocks[l].push_back(reg_vae_resnet("decoder.up_blocks." + std::to_string(l) + ".resnets." + std::to_string(i),
                                                   i == 0 ? c : co, co));
      c = co;
      if (l != L - 1) reg_conv("decoder.up_blocks." + std::to_string(l) + ".upsamplers.0.conv", c, c, 3);
    }
    reg_norm("decoder.conv_norm_out", c);
    reg_conv("decoder.conv_out", cfg.out_channels, c, 3);
  }
  void pack_vae_resnet(ResnetW& r) {
    r.g1 = rawf(r.key + ".norm1.weight"); r.b1 = rawf(r.key + ".norm1.bias");
    r.g2 = rawf(r.key + ".norm2.weight"); r.b2 = rawf(r.key + ".norm2.bias");
    r.conv1 = pack_conv(r.key + ".conv1", r.cout, r.cin, 3);
    if (r.shortcut) {  // conv2 and the 1x1 shortcut share one accumulator, as in the UNet
      r.conv2 = pack_conv(r.key + ".conv2", r.cout, r.cout, 3, r.cin);
      pack_rows(stream, rawf(r.key + ".conv_shortcut.weight"), r.cout, r.cin, r.cin, r.conv2.w, r.conv2.k, 9 * r.cout, nullptr);
      r.conv2.bias = bias_sum(r.key + ".conv2.bias", r.key + ".conv_shortcut.bias", r.cout);
    } else {
      r.conv2 = pack_conv(r.key + ".conv2", r.cout, r.cout, 3);
    }
  }
  void pack_vae_attn(VaeAttnW& a) {
    a.gn_g = rawf(a.key + ".group_norm.weight");
    a.gn_b = rawf(a.key + ".group_norm.bias");
    a.q = pack_linear(a.key + ".to_q", a.c, a.c, true);
    a.k = pack_linear(a.key + ".to_k", a.c, a.c, true);
    a.v = pack_linear(a.key + ".to_v", a.c, a.c, true);
    a.out = pack_linear(a.key + ".to_out.0", a.c, a.c, true);
  }
  std::vector<float> download(const std::string& k, size_t count) {
    std::vector<float> v(count);
    MV_CUDA(cudaMemcpyAsync(v.data(), rawf(k), count * sizeof(float), cudaMemcpyDeviceToHost, stream));
    MV_CUDA(cudaStreamSynchronize(stream));
    return v;
  }
  void finalize_vae(cudaStream_t s) {
    stream = s;
    for (auto& n : names) MV_CHECK(raw.count(n), "mvldm_finalize_weights: missing weight " + n);
    MV_CUDA(cudaStreamSynchronize(s));
    packed_store.clear();
    const int L = cfg.num_levels;
    const int* boc = cfg.block_out_channels;
    const int lc = cfg.latent_channels;
    auto pack_in = [&](const std::string& k, int cout, int cin) {  // conv on the explicit im2col operand, K padded to 64
      Packed p;
      p.n = cout;
      p.k = (9 * cin + 63) / 64 * 64;
      p.w = store<bf16>((size_t)p.n * p.k);
      pack_conv3x3(stream, rawf(k + ".weight"), cout, cin, 3, p.w, p.k, 0);
      p.bias = bias_sum(k + ".bias", "", cout);
      return p;
    };
    // ---- encoder
    vae_enc.conv_in = pack_in("encoder.conv_in", boc[0], cfg.in_channels);
    vae_enc.resample.assign(L, Packed{});
    for (int l = 0; l < L; ++l) {
      for (auto& r : vae_enc.blocks[l]) pack_vae_resnet(r);
      if (l != L - 1)
        vae_enc.resample[l] = pack_conv("encoder.down_blocks." + std::to_string(l) + ".downsamplers.0.conv", boc[l], boc[l], 3);
    }
    pack_vae_resnet(vae_enc.mid0);
    pack_vae_attn(vae_enc.attn);
    pack_vae_resnet(vae_enc.mid1);
    vae_enc.norm_g = rawf("encoder.conv_norm_out.weight");
    vae_enc.norm_b = rawf("encoder.conv_norm_out.bias");
    {  // quant_conv (1x1, 2*lc -> 2*lc) composed into conv_out: both are linear and nothing sits between them
      const int m2 = 2 * lc, cl = boc[L - 1], kk = 9 * cl;
      MV_CHECK(m2 <= 32, "latent_channels too large for the fp32 NCHW head (2 * latent_channels <= 32)");
      const std::vector<float> wc = download("encoder.conv_out.weight", (size_t)m2 * kk);  // [m2, cl, 3, 3]
      const std::vector<float> bc = download("encoder.conv_out.bias", m2);
      const std::vector<float> wq = download("quant_conv.weight", (size_t)m2 * m2);
      const std::vector<float> bq = download("quant_conv.bias", m2);
      std::vector<float> w2((size_t)m2 * kk, 0.f), b2(32, 0.f);
      for (int o = 0; o < m2; ++o) {
        double b = bq[o];
        for (int i = 0; i < m2; ++i) {
          b += (double)wq[o * m2 + i] * bc[i];
          const float f = wq[o * m2 + i];
          for (int e = 0; e < kk; ++e) w2[(size_t)o * kk + e] += f * wc[(size_t)i * kk + e];
        }
        b2[o] = (float)b;
      }
      float* dw = store<float>((size_t)m2 * kk);
      MV_CUDA(cudaMemcpyAsync(dw, w2.data(), w2.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
      vae_enc.conv_out.n = 32;
      vae_enc.conv_out.k = kk;
      vae_enc.conv_out.w = store<bf16>((size_t)32 * kk);
      pack_conv3x3(stream, dw, m2, cl, 3, vae_enc.conv_out.w, kk, 0);
      vae_enc.conv_out.bias = store<float>(32);
      MV_CUDA(cudaMemcpyAsync(vae_enc.conv_out.bias, b2.data(), 32 * sizeof(float), cudaMemcpyHostToDevice, stream));
      MV_CUDA(cudaStreamSynchronize(stream));
    }
    // ---- decoder
    {
      const std::vector<float> wq = download("post_quant_conv.weight", (size_t)lc * lc);
      const std::vector<float> bq = download("post_quant_conv.bias", lc);
      std::vector<float> pm(wq);
      pm.insert(pm.end(), bq.begin(), bq.end());
      float* d = store<float>(pm.size());
      MV_CUDA(cudaMemcpyAsync(d, pm.data(), pm.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
      MV_CUDA(cudaStreamSynchronize(stream));
      vae_premix = d;
    }
    vae_dec.conv_in = pack_in("decoder.conv_in", boc[L - 1], lc);
    pack_vae_resnet(vae_dec.mid0);
    pack_vae_attn(vae_dec.attn);
    pack_vae_resnet(vae_dec.mid1);
    vae_dec.resample.assign(L, Packed{});
    for (int l = 0; l < L; ++l) {
      for (auto& r : vae_dec.blocks[l]) pack_vae_resnet(r);
      if (l != L - 1)
        vae_dec.resample[l] = pack_conv("decoder.up_blocks." + std::to_string(l) + ".upsamplers.0.conv", boc[L - 1 - l],
                                        boc[L - 1 - l], 3);
    }
    vae_dec.norm_g = rawf("decoder.conv_norm_out.weight");
    vae_dec.norm_b = rawf("decoder.conv_norm_out.bias");
    vae_dec.conv_out = pack_conv("decoder.conv_out", cfg.out_channels, boc[0], 3, 0, 32);
    MV_CUDA(cudaStreamSynchronize(s));
    for (auto it = raw.begin(); it != raw.end();) {
      const bool keep = it->first.find("norm") != std::string::npos;
      it = keep ? std::next(it) : raw.erase(it);
    }
    plans.clear();
    prof_plan = last_plan = nullptr;
    finalized = true;
  }

  void im2col_step(const float* src, int n, int cin, int h, int w, int kpad, const float* premix, const Act& col) {
    if (dry) return;
    Step st;
    st.kind = Step::IM2COL;
    st.x0 = src; st.n_img = n; st.c0 = cin; st.h = h; st.w = w; st.aux = kpad; st.dst = col.p; st.gamma = premix;
    st.meta = OpMeta{"input_im2col", "", 0.0, (double)col.tokens() * (kpad * 2 + cin * 4)};
    push_step(st);
  }
  Act upsample2x(const Act& x) {
    Act u = new_act(x.n, x.h * 2, x.w * 2, x.c);
    if (!dry) {
      Step st;
      st.kind = Step::UPSAMPLE;
      st.x0 = x.p; st.n_img = x.n; st.h = x.h; st.w = x.w; st.c0 = x.c; st.dst = u.p;
      st.meta = OpMeta{"upsample", "", 0.0, 2.5 * (double)u.tokens() * u.c};
      push_step(st);
    }
    return u;
  }
  // head: GroupNorm -> SiLU -> conv3x3 -> fp32 NCHW
  void nchw_head(const Act& x, const float* g, const float* b, float eps, const Packed& w, int n_valid, float* out) {
    Act a = new_act(x.n, x.h, x.w, x.c);
    gn(x, nullptr, g, b, eps, true, a);
    mvldm_gemm_desc d{};
    d.nseg = 1;
    d.seg[0] = seg_conv3x3(a);
    d.n_img = x.n; d.oh = x.h; d.ow = x.w;
    d.w = w.w; d.n = w.n; d.k = w.k;
    d.bias = w.bias;
    d.mode = 2; d.out = out; d.n_valid = n_valid;
    run_gemm(d, 2.0 * (double)x.tokens() * n_valid * w.k);
  }
  // diffusers Attention(heads=1, dim_head=C, residual_connection, norm_num_groups, eps 1e-6) over the h*w tokens of each image.
  // One 512-wide head does not fit the flash kernels (head_dim <= 192), and at 1024 tokens per image the score matrix is 2 MB:
  // it runs as plain GEMMs on the tcgen05 kernel.  Per image: S = Q K^T (fp32) -> softmax rows (bf16) -> O = P V with V^T
  // produced directly by a GEMM whose "activation" operand is W_v and whose "weight" operand is the normalised input; b_v is
  // added after P V (softmax rows sum to 1).
  Act vae_attention(const VaeAttnW& a, const Act& x) {
    const int n = x.n, h = x.h, w = x.w, C = a.c, hw = h * w;
    MV_CHECK(hw % 64 == 0 && hw <= 4096, "VAE attention: tokens per image must be a multiple of 64, <= 4096");
    Act out = new_act(n, h, w, C);
    const size_t mark = arena.off;
    Act g = new_act(n, h, w, C);
    gn(x, nullptr, a.gn_g, a.gn_b, 1e-6f, false, g);
    Act q = new_act(n, h, w, C), k = new_act(n, h, w, C), o = new_act(n, h, w, C);
    gemm({seg_1x1(g)}, a.q, q);
    gemm({seg_1x1(g)}, a.k, k);
    float* S = new_f32((size_t)hw * hw);
    Act P = new_act(1, h, w, hw);       // [hw tokens, hw keys] bf16
    Act VT = new_act(1, C / 64, 64, hw);  // [C rows, hw] bf16
    Act wv;                             // W_v [C, C] viewed as an "image" of C pixels with C channels
    wv.n = 1; wv.h = C / 64; wv.w = 64; wv.c = C; wv.p = a.v.w;
    for (int i = 0; i < n; ++i) {
      Act qi = q, oi = o;
      qi.n = oi.n = 1;
      qi.p = q.p + (size_t)i * hw * C;
      oi.p = o.p + (size_t)i * hw * C;
      {  // S = Q_i K_i^T, fp32 row-major
        mvldm_gemm_desc d{};
        d.nseg = 1; d.seg[0] = seg_1x1(qi);
        d.n_img = 1; d.oh = h; d.ow = w;
        d.w = k.p + (size_t)i * hw * C; d.n = hw; d.k = C;
        d.mode = 4; d.out = S; d.ldo = hw; d.n_valid = hw;
        run_gemm(d);
      }
      if (!dry) {
        Step st;
        st.kind = Step::SOFTMAX;
        st.x0 = S; st.n_img = hw; st.c0 = hw; st.eps = 1.f / sqrtf((float)C); st.dst = P.p;
        st.meta = OpMeta{"softmax", "rows" + std::to_string(hw), 0.0, 6.0 * (double)hw * hw};
        push_step(st);
      }
      {  // V^T = W_v g_i^T   ([C, hw]; b_v is added after P V)
        mvldm_gemm_desc d{};
        d.nseg = 1; d.seg[0] = seg_1x1(wv);
        d.n_img = 1; d.oh = wv.h; d.ow = wv.w;
        d.w = g.p + (size_t)i * hw * C; d.n = hw; d.k = C;
        d.mode = 0; d.out = VT.p; d.ldo = hw; d.n_valid = hw;
        run_gemm(d);
      }
      {  // O_i = P V + b_v
        mvldm_gemm_desc d{};
        d.nseg = 1; d.seg[0] = seg_1x1(P);
        d.n_img = 1; d.oh = h; d.ow = w;
        d.w = VT.p; d.n = C; d.k = hw;
        d.bias = a.v.bias;
        d.mode = 0; d.out = oi.p; d.ldo = C; d.n_valid = C;
        run_gemm(d);
      }
    }
    gemm({seg_1x1(o)}, a.out, out, nullptr, 0, &x);
    if (!taps_enabled) arena.off = mark;
    return out;
  }
  // AutoencoderKL.decode(z).sample  (diffusion_wrapper.py:293-295; z already divided by the scaling factor by the caller)
  void run_vae_decode(const float* z, int n, int Hh, int Ww, float* image) {
    const int L = cfg.num_levels;
    arena.off = 0;
    const int kpad = vae_dec.conv_in.k;
    Act col = new_act(n, Hh, Ww, kpad);
    im2col_step(z, n, cfg.latent_channels, Hh, Ww, kpad, vae_premix, col);
    Act x = new_act(n, Hh, Ww, vae_dec.conv_in.n);
    gemm({seg_1x1(col)}, vae_dec.conv_in, x);
    tap("decoder.conv_in", x);
    x = resnet(vae_dec.mid0, x, nullptr, nullptr);
    x = vae_attention(vae_dec.attn, x);
    x = resnet(vae_dec.mid1, x, nullptr, nullptr);
    tap("decoder.mid", x);
    for (int l = 0; l < L; ++l) {
      for (auto& r : vae_dec.blocks[l]) x = resnet(r, x, nullptr, nullptr);
      if (l != L - 1) {
        Act u = upsample2x(x);
        Act y = new_act(n, u.h, u.w, u.c);
        gemm({seg_conv3x3(u)}, vae_dec.resample[l], y);
        x = y;
      }
      tap("decoder.up" + std::to_string(l), x);
    }
    nchw_head(x, vae_dec.norm_g, vae_dec.norm_b, 1e-6f, vae_dec.conv_out, cfg.out_channels, image);
  }
  // AutoencoderKL.encode(x).latent_dist.parameters: [n, 2*latent, H/f, W/f] = (mean | logvar)  (diffusion_wrapper.py:283)
  void run_vae_encode(const float* image, int n, int Hh, int Ww, float* moments) {
    const int L = cfg.num_levels;
    arena.off = 0;
    const int kpad = vae_enc.conv_in.k;
    Act col = new_act(n, Hh, Ww, kpad);
    im2col_step(image, n, cfg.in_channels, Hh, Ww, kpad, nullptr, col);
    Act x = new_act(n, Hh, Ww, vae_enc.conv_in.n);
    gemm({seg_1x1(col)}, vae_enc.conv_in, x);
    tap("encoder.conv_in", x);
    for (int l = 0; l < L; ++l) {
      for (auto& r : vae_enc.blocks[l]) x = resnet(r, x, nullptr, nullptr);
      if (l != L - 1) {
        // Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) then conv3x3 stride 2 -> taps at +0, +1, +2 from (2y, 2x); the
        // pad row / column is the TMA's out-of-bounds zero fill
        Act y = new_act(n, x.h / 2, x.w / 2, x.c);
        mvldm_aseg sg = seg_conv3x3(x, 2);
        for (int t = 0; t < 9; ++t) {
          sg.dh[t] = (int8_t)(t / 3);
          sg.dw[t] = (int8_t)(t % 3);
        }
        gemm({sg}, vae_enc.resample[l], y);
        x = y;
      }
      tap("encoder.down" + std::to_string(l), x);
    }
    x = resnet(vae_enc.mid0, x, nullptr, nullptr);
    x = vae_attention(vae_enc.attn, x);
    x = resnet(vae_enc.mid1, x, nullptr, nullptr);
    tap("encoder.mid", x);
    nchw_head(x, vae_enc.norm_g, vae_enc.norm_b, 1e-6f, vae_enc.conv_out, 2 * cfg.latent_channels, moments);
  }

  std::vector<int> scene_views;  // of the forward being recorded / run
  static int total_views(const std::vector<int>& sv) {
    int n = 0;
    for (int v : sv) n += v;
    return n;
  }

  void run(const float* latents, const int64_t* tsteps, const std::vector<int>& sv, int Hh, int Ww, float* out_eps) {
    scene_views = sv;
    const int L = cfg.num_levels, n = total_views(sv);
    const int* boc = cfg.block_out_channels;
    arena.off = 0;
    // ---- time embedding (K2): sinusoid -> linear -> SiLU -> linear -> SiLU -> all time_emb_proj of the net at once.
    // Three tensor-core GEMMs over the n (<= 128 per tile) rows: the SIMT row-by-column kernel this replaces re-read
    // the activations once per output column (90 us at 8 views, 560 us at 64).
    Act sinus = new_act(n, 1, 1, boc[0]);
    Act e1 = new_act(n, 1, 1, temb_dim);
    Act e2 = new_act(n, 1, 1, temb_dim);
    float* temb = new_f32((size_t)n * temb_total);
    const size_t temb_first = dry ? 0 : rec->steps.size();
    if (!dry) {
      Step st;
      st.kind = Step::SINUSOID;
      st.x0 = tsteps; st.n_img = n; st.c0 = boc[0]; st.dst = sinus.p;
      st.meta = OpMeta{"time_embedding", "sinusoid rows" + std::to_string(n), 0.0, 0.0};
      push_step(st);
    }
    gemm({seg_1x1(sinus)}, time1, e1, nullptr, 0, nullptr, 3);
    gemm({seg_1x1(e1)}, time2, e2, nullptr, 0, nullptr, 3);  // SiLU(emb): every consumer (ResnetBlock2D) applies it first
    {
      mvldm_gemm_desc d{};
      d.nseg = 1;
      d.seg[0] = seg_1x1(e2);
      d.n_img = n; d.oh = 1; d.ow = 1;
      d.w = temb_all.w; d.n = temb_all.n; d.k = temb_all.k;
      d.bias = temb_all.bias;
      d.mode = 4; d.out = temb; d.ldo = temb_total; d.n_valid = temb_all.n;
      run_gemm(d);
    }
    if (!dry)  // (none of these GEMMs is split-K - modes 3 / 4 never are - so they do not touch the shared split-K scratch)
      for (size_t i = temb_first; i < rec->steps.size(); ++i) rec->steps[i].lane = 1;
    // ---- conv_in on the im2col'd fp32 input
    Act col = new_act(n, Hh, Ww, kpad_in);
    if (!dry) {
      Step st;
      st.kind = Step::IM2COL;
      st.x0 = latents; st.n_img = n; st.c0 = cfg.in_channels; st.h = Hh; st.w = Ww; st.aux = kpad_in; st.dst = col.p;
      st.meta = OpMeta{"input_im2col", "", 0.0, (double)col.tokens() * (kpad_in * 2 + cfg.in_channels * 4)};
      push_step(st);
    }
    Act x = new_act(n, Hh, Ww, boc[0]);
    gemm({seg_1x1(col)}, conv_in, x, nullptr, 0, nullptr, 0, 2.0 * (double)x.tokens() * boc[0] * 9 * cfg.in_channels);
    tap("conv_in", x);
    std::vector<Act> skips{x};
    // ---- down
    for (int l = 0; l < L; ++l) {
      for (int i = 0; i < cfg.layers_per_block; ++i) {
        x = resnet(down_res[l][i], x, nullptr, temb);
        if (!t2d_down[l].empty()) x = t2d_block(t2d_down[l][i], x);
        tap("down" + std::to_string(l) + ".res" + std::to_string(i), x);
        skips.push_back(x);
      }
      if (x.h <= cfg.max_attn_res && x.w <= cfg.max_attn_res) {
        x = mv_block(mv_enc[l], x);
        tap("down" + std::to_string(l) + ".mv", x);
      }
      if (l != L - 1) {
        Act y = new_act(n, x.h / 2, x.w / 2, x.c);
        gemm({seg_conv3x3(x, 2)}, down_conv[l], y);
        x = y;
        tap("down" + std::to_string(l) + ".ds", x);
        skips.push_back(x);
      }
    }
    // ---- mid
    x = resnet(mid_res, x, nullptr, temb);
    if (cfg.variant == 1) {
      x = t2d_block(t2d_mid, x);
      x = resnet(mid_res1, x, nullptr, temb);
    }
    tap("mid.res0", x);
    x = mv_block(mv_mid, x);
    tap("mid.mv", x);
    // ---- up
    for (int l = 0; l < L; ++l) {
      for (int i = 0; i < cfg.layers_per_block + 1; ++i) {
        Act skip = skips.back();
        skips.pop_back();
        x = resnet(up_res[l][i], x, &skip, temb);
        tap("up" + std::to_string(l) + ".res" + std::to_string(i), x);
      }
      if (x.h <= cfg.max_attn_res && x.w <= cfg.max_attn_res) {
        x = mv_block(mv_dec[l], x);
        tap("up" + std::to_string(l) + ".mv", x);
      }
      if (l != L - 1) {
        Act u = new_act(n, x.h * 2, x.w * 2, x.c);
        if (!dry) {
          Step st;
          st.kind = Step::UPSAMPLE;
          st.x0 = x.p; st.n_img = n; st.h = x.h; st.w = x.w; st.c0 = x.c; st.dst = u.p;
          st.meta = OpMeta{"upsample", "", 0.0, 2.5 * (double)u.tokens() * u.c};
          push_step(st);
        }
        Act y = new_act(n, u.h, u.w, u.c);
        gemm({seg_conv3x3(u)}, up_conv[l], y);
        x = y;
        tap("up" + std::to_string(l) + ".us", x);
      }
    }
    // ---- head: GN -> SiLU -> conv3x3 -> fp32 NCHW
    Act a = new_act(n, x.h, x.w, x.c);
    gn(x, nullptr, norm_out_g, norm_out_b, 1e-5f, true, a);
    mvldm_gemm_desc d{};
    d.nseg = 1;
    d.seg[0] = seg_conv3x3(a);
    d.n_img = n; d.oh = x.h; d.ow = x.w;
    d.w = conv_out.w; d.n = conv_out.n; d.k = conv_out.k;
    d.bias = conv_out.bias;
    d.mode = 2; d.out = out_eps; d.n_valid = cfg.out_channels;
    run_gemm(d, 2.0 * (double)x.tokens() * cfg.out_channels * conv_out.k);
  }

  // bytes of the fp32 NCHW tensor a program reads (in) / writes (out) for n images of H x W input pixels
  size_t io_bytes(int n, int H, int W, bool in) const {
    const int f = 1 << (cfg.num_levels - 1);
    size_t per = 0;
    if (program == 0) per = (size_t)(in ? cfg.in_channels : cfg.out_channels) * H * W;
    else if (program == 1) per = in ? (size_t)cfg.latent_channels * H * W : (size_t)cfg.out_channels * H * f * W * f;
    else per = in ? (size_t)cfg.in_channels * H * W : (size_t)2 * cfg.latent_channels * (H / f) * (W / f);
    return (size_t)n * per * sizeof(float);
  }
  void run_program(const float* in, const int64_t* tsteps, const std::vector<int>& sv, int H, int W, float* out) {
    if (program == 0) run(in, tsteps, sv, H, W, out);
    else if (program == 1) run_vae_decode(in, total_views(sv), H, W, out);
    else run_vae_encode(in, total_views(sv), H, W, out);
  }

  // ---- plans: measured, allocated and recorded once per batch shape; a small LRU keeps the device memory bounded ----
  static constexpr size_t kMaxPlans = 6;
  Plan& plan_for(const std::vector<int>& sv, int H, int W, bool shard = false) {
    std::vector<int> key{H, W, taps_enabled ? 1 : 0, shard ? v_total : 0, program};
    if (shard) {  // the recorded launch list holds the caller's exchange buffers
      const uint64_t a = reinterpret_cast<uint64_t>(kv_send), b = reinterpret_cast<uint64_t>(kv_recv);
      for (uint64_t v : {a, b}) {
        key.push_back((int)(v & 0xffffffffu));
        key.push_back((int)(v >> 32));
      }
    }
    key.insert(key.end(), sv.begin(), sv.end());
    const int n = total_views(sv);
    auto it = plans.find(key);
    if (it != plans.end()) {
      it->second->last_use = ++use_clock;
      return *it->second;
    }
    if (plans.size() >= kMaxPlans) {  // evict the least recently used plan (arena, scratch, graph)
      MV_CUDA(cudaDeviceSynchronize());
      auto victim = plans.begin();
      for (auto jt = plans.begin(); jt != plans.end(); ++jt)
        if (jt->second->last_use < victim->second->last_use) victim = jt;
      if (prof_plan == victim->second.get()) prof_plan = nullptr;
      if (last_plan == victim->second.get()) last_plan = nullptr;
      plans.erase(victim);
    }
    groupnorm_init();
    std::unique_ptr<Plan> p(new Plan());
    p->scene_views = sv; p->H = H; p->W = W;
    p->last_use = ++use_clock;
    sharded = shard;
    // pass 1: arena offsets and split-K scratch
    dry = true;
    arena = Arena();
    arena.measuring = true;
    splitk_need = 0;
    run_program(nullptr, nullptr, sv, H, W, nullptr);
    p->arena_bytes = arena.peak;
    p->arena_mem.alloc(p->arena_bytes);
    p->splitk.alloc(splitk_need);
    if (splitk_need) MV_CUDA(cudaMemset(p->splitk.p, 0, splitk_need));  // fused split-K counters start (and end) at zero
    p->in_latents.alloc(io_bytes(n, H, W, true));
    p->in_t.alloc((size_t)n * sizeof(int64_t));
    p->out_eps.alloc(io_bytes(n, H, W, false));
    // pass 2: record the launch list against the real buffers
    dry = false;
    arena = Arena();
    arena.measuring = false;
    arena.base = reinterpret_cast<char*>(p->arena_mem.p);
    arena.cap = p->arena_bytes;
    splitk_ws = p->splitk.p;
    splitk_bytes = p->splitk.bytes;
    rec = p.get();
    try {
      run_program((const float*)p->in_latents.p, (const int64_t*)p->in_t.p, sv, H, W, (float*)p->out_eps.p);
    } catch (...) {
      rec = nullptr;
      sharded = false;
      throw;
    }
    rec = nullptr;
    sharded = false;
    p->launches = (int)p->steps.size();
    Plan& ref = *p;
    plans[key] = std::move(p);
    return ref;
  }

  // issue the plan's launches on `s` (eagerly, or into a stream capture); with `events` every launch is bracketed
  void execute(Plan& p, cudaStream_t main_s, std::vector<cudaEvent_t>* events) {
    size_t ev = 0;
    // lanes: under the per-op profiler everything stays on one stream (event pairs must bracket single launches)
    static const bool lanes_on = [] {
      const char* e = getenv("MVLDM_LANES");
      return !e || atoi(e) != 0;
    }();
    const bool lanes = lanes_on && !events;
    if (lanes && !side_stream) {
      MV_CUDA(cudaStreamCreateWithFlags(&side_stream, cudaStreamNonBlocking));
      MV_CUDA(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
      MV_CUDA(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
    }
    bool forked = false;
    auto join = [&] {
      if (!forked) return;
      MV_CUDA(cudaEventRecord(ev_join, side_stream));
      MV_CUDA(cudaStreamWaitEvent(main_s, ev_join, 0));
      forked = false;
    };
    for (const Step& st : p.steps) {
      cudaStream_t s = main_s;
      if (lanes && st.lane == 1) {
        if (!forked) {  // the side lane starts after everything already enqueued on the main one (input staging copies)
          MV_CUDA(cudaEventRecord(ev_fork, main_s));
          MV_CUDA(cudaStreamWaitEvent(side_stream, ev_fork, 0));
          forked = true;
        }
        s = side_stream;
      } else if (st.join_side) {
        join();
      }
      if (events) MV_CUDA(cudaEventRecord((*events)[ev++], s));
      switch (st.kind) {
        case Step::GEMM:
          gemm_tc(s, st.gemm, p.splitk.p, p.splitk.bytes);
          break;
        case Step::GEMM_SIMT:
          gemm_simt(s, st.gemm);
          break;
        case Step::ATTN:
          if (st.simt) attention_simt(s, st.qkv, st.out, st.batches, st.seq, st.heads, st.d, st.dpad);
          else attention_tc(s, st.qkv, st.out, st.batches, st.seq, st.heads, st.d, st.dpad);
          break;
        case Step::ATTN_SHARDED: {
          const int hd = st.heads * st.dpad;
          const int world = st.seq / st.seq_local, g = group_index;
          const size_t bytes = (size_t)st.seq_local * 2 * hd * sizeof(bf16);
          MV_CHECK(bytes * world <= kv_recv_bytes, "mvldm_forward_sharded: kv_recv buffer too small");
          // K|V columns of the packed q|k|v rows -> contiguous slab
          MV_CUDA(cudaMemcpy2DAsync(kv_send, (size_t)2 * hd * sizeof(bf16), st.qkv + hd, (size_t)3 * hd * sizeof(bf16),
                                    (size_t)2 * hd * sizeof(bf16), st.seq_local, cudaMemcpyDeviceToDevice, s));
          int rc = exchange(exchange_user, kv_send, kv_recv, (int64_t)bytes, s, MVLDM_EXCHANGE_BEGIN);
          MV_CHECK(rc == 0 || rc == MVLDM_EXCHANGE_DONE, "mvldm_forward_sharded: K/V exchange callback failed");
          const bf16* own = reinterpret_cast<const bf16*>(kv_send);
          const bf16* all = reinterpret_cast<const bf16*>(kv_recv);
          if (rc == MVLDM_EXCHANGE_DONE && world > 1) {
            // the host ran the whole exchange on `s`: nothing to overlap, one pass over all slabs in view order (the same
            // softmax order on every rank, and the one a single GPU would use)
            attention_tc_kv(s, st.qkv, 3 * hd, 0, all, 2 * hd, 0, hd, st.out, 1, st.seq_local, st.seq, st.heads, st.d, st.dpad);
            break;
          }
          if (world == 1) {
            attention_tc_kv(s, st.qkv, 3 * hd, 0, own, 2 * hd, 0, hd, st.out, 1, st.seq_local, st.seq_local, st.heads, st.d,
                            st.dpad);
            rc = exchange(exchange_user, kv_send, kv_recv, (int64_t)bytes, s, MVLDM_EXCHANGE_END);
            MV_CHECK(rc == 0, "mvldm_forward_sharded: K/V exchange callback failed");
            break;
          }
          // this rank's queries against its OWN keys run while the other ranks' slabs are still on the wire ...
          const bf16* parts[3];
          const float* stats[3];
          int np = 0;
          attention_tc_kv(s, st.qkv, 3 * hd, 0, own, 2 * hd, 0, hd, st.part[0], 1, st.seq_local, st.seq_local, st.heads, st.d,
                          st.dpad, st.pstats[0]);
          parts[np] = st.part[0]; stats[np++] = st.pstats[0];
          rc = exchange(exchange_user, kv_send, kv_recv, (int64_t)bytes, s, MVLDM_EXCHANGE_END);
          MV_CHECK(rc == 0, "mvldm_forward_sharded: K/V exchange callback failed");
          // ... then against the slabs in front of and behind its own, and the three softmax states are merged in this
          // fixed order (deterministic; equal to the one-pass softmax up to fp32 rounding of the merge)
          if (g > 0) {
            attention_tc_kv(s, st.qkv, 3 * hd, 0, all, 2 * hd, 0, hd, st.part[1], 1, st.seq_local, g * st.seq_local, st.heads,
                            st.d, st.dpad, st.pstats[1]);
            parts[np] = st.part[1]; stats[np++] = st.pstats[1];
          }
          if (g < world - 1) {
            attention_tc_kv(s, st.qkv, 3 * hd, 0, all + (size_t)(g + 1) * st.seq_local * 2 * hd, 2 * hd, 0, hd, st.part[2], 1,
                            st.seq_local, (world - 1 - g) * st.seq_local, st.heads, st.d, st.dpad, st.pstats[2]);
            parts[np] = st.part[2]; stats[np++] = st.pstats[2];
          }
          attention_merge(s, np, parts, stats, st.seq_local, st.heads, st.dpad, st.out);
          break;
        }
        case Step::GN:
          groupnorm(s, (const bf16*)st.x0, st.c0, (const bf16*)st.x1, st.c1, st.n_img, st.hw, st.groups, st.eps, st.gamma,
                    st.beta, st.silu != 0, (bf16*)st.dst, st.scratch);
          break;
        case Step::LN:
          layernorm(s, (const bf16*)st.x0, st.n_img, st.c0, st.eps, st.gamma, st.beta, (bf16*)st.dst);
          break;
        case Step::UPSAMPLE:
          upsample_nearest2x(s, (const bf16*)st.x0, st.n_img, st.h, st.w, st.c0, (bf16*)st.dst);
          break;
        case Step::IM2COL:
          im2col_input(s, (const float*)st.x0, st.n_img, st.c0, st.h, st.w, st.aux, (bf16*)st.dst, st.gamma);
          break;
        case Step::SOFTMAX:
          softmax_rows(s, (const float*)st.x0, st.n_img, st.c0, st.eps, (bf16*)st.dst);
          break;
        case Step::SINUSOID:
          timestep_sinusoid_bf16(s, (const int64_t*)st.x0, st.n_img, st.c0, (bf16*)st.dst);
          break;
      }
      if (events) MV_CUDA(cudaEventRecord((*events)[ev++], s));
    }
    join();  // (a plan whose side lane nobody read)
  }

  void forward(cudaStream_t s, const float* latents, const int64_t* tsteps, const std::vector<int>& sv, int H, int W,
               float* out, int prog = 0) {
    MV_CHECK(finalized, "mvldm_forward before mvldm_finalize_weights");
    MV_CHECK(!sv.empty(), "empty batch");
    for (int v : sv) MV_CHECK(v > 0, "empty scene");
    MV_CHECK((cfg.model == MVLDM_MODEL_VAE) == (prog != 0), "entry point does not match the handle's model (denoiser / VAE)");
    const int n = total_views(sv);
    const int down = 1 << (cfg.num_levels - 1);
    if (prog != 1) MV_CHECK(H % down == 0 && W % down == 0, "input size must be divisible by 2^(levels-1)");
    stream = s;
    program = prog;
    Plan& p = plan_for(sv, H, W);
    last_plan = &p;
    const size_t in_bytes = io_bytes(n, H, W, true);
    const size_t out_bytes = io_bytes(n, H, W, false);
    // the recorded launches read and write fixed buffers, so the list (and its graph) is valid for any caller tensors
    MV_CUDA(cudaMemcpyAsync(p.in_latents.p, latents, in_bytes, cudaMemcpyDeviceToDevice, s));
    if (tsteps) MV_CUDA(cudaMemcpyAsync(p.in_t.p, tsteps, (size_t)n * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    const bool graph = cfg.use_cuda_graph && !taps_enabled && !profiling;
    cudaStreamCaptureStatus cst;
    MV_CUDA(cudaStreamIsCapturing(s, &cst));
    if (!graph || cst != cudaStreamCaptureStatusNone) {  // eager, or the caller is capturing: record into their graph
      if (profiling && cst == cudaStreamCaptureStatusNone) {
        while (event_pool.size() < 2 * p.steps.size()) {
          cudaEvent_t e;
          MV_CUDA(cudaEventCreate(&e));
          event_pool.push_back(e);
        }
        g_launch_count = 0;
        execute(p, s, &event_pool);
        p.launches = g_launch_count;
        prof_plan = &p;
      } else {
        g_launch_count = 0;
        execute(p, s, nullptr);
        p.launches = g_launch_count;
      }
    } else {
      if (!p.graph) {
        cudaGraph_t g = nullptr;
        if (!capture_stream) MV_CUDA(cudaStreamCreateWithFlags(&capture_stream, cudaStreamNonBlocking));
        MV_CUDA(cudaStreamBeginCapture(capture_stream, cudaStreamCaptureModeThreadLocal));
        try {
          g_launch_count = 0;
          execute(p, capture_stream, nullptr);
          p.launches = g_launch_count;
        } catch (...) {
          cudaStreamEndCapture(capture_stream, &g);
          if (g) cudaGraphDestroy(g);
          throw;
        }
        MV_CUDA(cudaStreamEndCapture(capture_stream, &g));
        cudaError_t e = cudaGraphInstantiate(&p.graph, g, 0);
        cudaGraphDestroy(g);
        MV_CUDA(e);
      }
      MV_CUDA(cudaGraphLaunch(p.graph, s));
    }
    last_launches = p.launches;
    MV_CUDA(cudaMemcpyAsync(out, p.out_eps.p, out_bytes, cudaMemcpyDeviceToDevice, s));
  }
};

// Folds the event pairs of the last profiled forward into a JSON string (synchronises the device).
static const char* profile_report(mvldm_handle_s* h) {
  MV_CUDA(cudaDeviceSynchronize());
  MV_CHECK(h->prof_plan != nullptr, "mvldm_profile_json: run a forward with profiling enabled first");
  Plan& p = *h->prof_plan;
  struct Agg {
    int n = 0;
    double us = 0, flops = 0, bytes = 0, scores = 0, flops_padded = 0;
  };
  std::map<std::string, Agg> agg;
  std::string ops = "[";
  size_t ev = 0;
  for (size_t i = 0; i < p.steps.size(); ++i) {
    const OpMeta& m = p.steps[i].meta;
    float ms = 0.f;
    MV_CUDA(cudaEventElapsedTime(&ms, h->event_pool[ev], h->event_pool[ev + 1]));
    ev += 2;
    Agg& a = agg[m.cat];
    a.n++; a.us += ms * 1e3; a.flops += m.flops; a.bytes += m.bytes; a.scores += m.scores; a.flops_padded += m.flops_padded;
    char buf[320];
    snprintf(buf, sizeof buf, "%s{\"cat\":\"%s\",\"what\":\"%s\",\"us\":%.2f,\"gflop\":%.3f}", i ? "," : "", m.cat,
             m.what.c_str(), ms * 1e3, m.flops * 1e-9);
    ops += buf;
  }
  ops += "]";
  std::string out = "{\"categories\":{";
  bool first = true;
  for (auto& kv : agg) {
    char buf[384];
    snprintf(buf, sizeof buf,
             "%s\"%s\":{\"launches\":%d,\"us\":%.2f,\"gflop\":%.3f,\"mbytes\":%.3f,\"gscores\":%.4f,\"gflop_padded\":%.3f}",
             first ? "" : ",", kv.first.c_str(), kv.second.n, kv.second.us, kv.second.flops * 1e-9, kv.second.bytes * 1e-6,
             kv.second.scores * 1e-9, kv.second.flops_padded * 1e-9);
    out += buf;
    first = false;
  }
  out += "},\"ops\":" + ops + "}";
  h->prof_json = out;
  return h->prof_json.c_str();
}

// =============================================================================================
// C ABI
// =============================================================================================
#define MV_API_BEGIN try {
#define MV_API_END                           \
  }                                          \
  catch (const std::exception& e) {          \
    mvldm::g_last_error = e.what();          \
    return 1;                                \
  }                                          \
  catch (...) {                              \
    mvldm::g_last_error = "unknown error";   \
    return 1;                                \
  }                                          \
  return 0;

extern "C" {

const char* mvldm_last_error(void) { return mvldm::g_last_error.c_str(); }
int mvldm_version(void) { return 100; }

int mvldm_create(const mvldm_config* cfg, int device, mvldm_handle* out) {
  MV_API_BEGIN
  MV_CHECK(cfg && out, "null argument");
  MV_CHECK(cfg->num_levels >= 1 && cfg->num_levels <= MVLDM_MAX_LEVELS, "num_levels out of range");
  int ndev = 0;
  MV_CUDA(cudaGetDeviceCount(&ndev));
  MV_CHECK(device >= 0 && device < ndev, "no such CUDA device (mvldm_b200 has no CPU fallback)");
  MV_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  MV_CUDA(cudaGetDeviceProperties(&prop, device));
  MV_CHECK(prop.major == 10, "mvldm_b200 is built for sm_100a (B200) only; found sm_" + std::to_string(prop.major) +
                                 std::to_string(prop.minor));
  std::unique_ptr<mvldm_handle_s> h(new mvldm_handle_s());
  h->cfg = *cfg;
  h->device = device;
  for (int l = 0; l < cfg->num_levels; ++l) {
    MV_CHECK(cfg->block_out_channels[l] % 64 == 0, "block_out_channels must be multiples of 64");
    MV_CHECK(cfg->block_out_channels[l] % cfg->norm_groups == 0, "channels not divisible by norm_groups");
  }
  MV_CHECK(cfg->variant == 0 || cfg->variant == 1, "variant must be 0 (A) or 1 (B)");
  if (cfg->variant == 1) MV_CHECK(cfg->cross_attention_dim > 0, "variant B needs cross_attention_dim");
  MV_CHECK(cfg->mv_block == MVLDM_MV_SPATIAL_TRANSFORMER_3D || cfg->mv_block == MVLDM_MV_STANDARD,
           "mv_block must be MVLDM_MV_SPATIAL_TRANSFORMER_3D or MVLDM_MV_STANDARD");
  if (cfg->mv_block == MVLDM_MV_STANDARD) {
    MV_CHECK(cfg->mv_num_layers >= 1 && cfg->mv_num_layers <= 8, "standard transformer: num_layers must be in [1, 8]");
    MV_CHECK((cfg->mv_d_mlp > 0) != (cfg->mv_d_mlp_multiplier > 0),
             "standard transformer: exactly one of d_mlp and d_mlp_multiplier (standard/transformer.py:63-64)");
  }
  MV_CHECK(cfg->model == MVLDM_MODEL_DENOISER || cfg->model == MVLDM_MODEL_VAE, "model must be MVLDM_MODEL_DENOISER or _VAE");
  if (cfg->model == MVLDM_MODEL_VAE) {
    MV_CHECK(cfg->latent_channels >= 1 && cfg->latent_channels <= 8, "VAE: latent_channels must be in [1, 8]");
    MV_CHECK(cfg->out_channels >= 1 && cfg->out_channels <= 32 && cfg->in_channels >= 1 && cfg->in_channels <= 7,
             "VAE: image channels out of range");
    MV_CHECK(cfg->layers_per_block >= 1, "VAE: layers_per_block must be >= 1");
    MV_CHECK(cfg->impl == MVLDM_IMPL_TC, "VAE: the tcgen05 kernels only");
    h->build_vae_registry();
  } else {
    h->build_registry();
  }
  *out = h.release();
  MV_API_END
}

int mvldm_destroy(mvldm_handle h) {
  MV_API_BEGIN
  if (h) {
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->capture_stream) cudaStreamDestroy(h->capture_stream);
    if (h->side_stream) cudaStreamDestroy(h->side_stream);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    for (cudaEvent_t e : h->event_pool) cudaEventDestroy(e);
    delete h;
  }
  MV_API_END
}

int mvldm_num_weights(mvldm_handle h) { return h ? (int)h->names.size() : -1; }
const char* mvldm_weight_name(mvldm_handle h, int i) {
  return (h && i >= 0 && i < (int)h->names.size()) ? h->names[i].c_str() : nullptr;
}
int mvldm_weight_shape(mvldm_handle h, int i, int64_t shape[4], int* ndim) {
  MV_API_BEGIN
  MV_CHECK(h && i >= 0 && i < (int)h->names.size(), "bad weight index");
  const auto& s = h->shapes[h->names[i]];
  *ndim = (int)s.size();
  for (size_t j = 0; j < s.size(); ++j) shape[j] = s[j];
  MV_API_END
}

int mvldm_set_weight(mvldm_handle h, const char* key, const void* ptr, const int64_t* shape, int ndim, int dtype,
                     void* stream) {
  MV_API_BEGIN
  MV_CHECK(h && key && ptr && shape, "null argument");
  MV_CUDA(cudaSetDevice(h->device));
  auto it = h->shapes.find(key);
  MV_CHECK(it != h->shapes.end(), std::string("unknown weight key: ") + key);
  MV_CHECK((int)it->second.size() == ndim, std::string("rank mismatch for ") + key);
  int64_t numel = 1;
  for (int i = 0; i < ndim; ++i) {
    MV_CHECK(it->second[i] == shape[i], std::string("shape mismatch for ") + key);
    numel *= shape[i];
  }
  MV_CHECK(dtype >= MVLDM_F32 && dtype <= MVLDM_F16, "bad dtype");
  if (h->unused.count(key)) return 0;  // shape-checked, never read by the forward: not kept on the device
  auto& buf = h->raw[key];
  if (!buf || buf->bytes != (size_t)numel * sizeof(float)) buf.reset(new DevBuf((size_t)numel * sizeof(float)));
  convert_f32((cudaStream_t)stream, ptr, dtype, numel, reinterpret_cast<float*>(buf->p));
  h->finalized = false;
  MV_API_END
}

int mvldm_finalize_weights(mvldm_handle h, void* stream) {
  MV_API_BEGIN
  MV_CHECK(h, "null handle");
  MV_CUDA(cudaSetDevice(h->device));
  if (h->cfg.model == MVLDM_MODEL_VAE) h->finalize_vae((cudaStream_t)stream);
  else h->finalize((cudaStream_t)stream);
  MV_API_END
}

int64_t mvldm_workspace_bytes(mvldm_handle h, int B, int V, int H, int W) {
  try {
    MV_CHECK(h && h->finalized, "finalize weights first");
    MV_CHECK(B > 0 && V > 0, "empty batch");
    MV_CHECK(h->cfg.model == MVLDM_MODEL_DENOISER, "mvldm_workspace_bytes: denoiser handles only");
    h->program = 0;
    return (int64_t)h->plan_for(std::vector<int>(B, V), H, W).arena_bytes;
  } catch (const std::exception& e) {
    mvldm::g_last_error = e.what();
    return -1;
  }
}

int mvldm_forward(mvldm_handle h, void* stream, const float* latents, const int64_t* timesteps, int B, int V, int H,
                  int W, float* out) {
  MV_API_BEGIN
  MV_CHECK(h && latents && timesteps && out, "null argument");
  MV_CUDA(cudaSetDevice(h->device));
  MV_CHECK(B > 0 && V > 0, "empty batch");
  h->forward((cudaStream_t)stream, latents, timesteps, std::vector<int>(B, V), H, W, out);
  MV_API_END
}

int mvldm_forward_scenes(mvldm_handle h, void* stream, const float* latents, const int64_t* timesteps, int num_scenes,
                         const int32_t* views_per_scene, int H, int W, float* out) {
  MV_API_BEGIN
  MV_CHECK(h && latents && timesteps && out && views_per_scene, "null argument");
  MV_CHECK(num_scenes > 0 && num_scenes <= 4096, "num_scenes out of range");
  MV_CUDA(cudaSetDevice(h->device));
  h->forward((cudaStream_t)stream, latents, timesteps, std::vector<int>(views_per_scene, views_per_scene + num_scenes), H, W,
             out);
  MV_API_END
}

int mvldm_forward_sharded(mvldm_handle h, void* stream, const float* latents, const int64_t* timesteps, int V_local,
                          int V_total, int group_index, int H, int W, float* out, void* kv_send, void* kv_recv,
                          int64_t kv_recv_bytes, mvldm_kv_exchange_fn exchange, void* user) {
  MV_API_BEGIN
  MV_CHECK(h && latents && timesteps && out && kv_send && kv_recv && exchange, "null argument");
  MV_CHECK(h->finalized, "mvldm_forward_sharded before mvldm_finalize_weights");
  MV_CHECK(h->cfg.impl == MVLDM_IMPL_TC, "view-group sharding needs the tcgen05 kernels");
  MV_CHECK(h->cfg.model == MVLDM_MODEL_DENOISER, "mvldm_forward_sharded: denoiser handles only");
  MV_CHECK(V_local > 0 && V_total % V_local == 0, "V_total must be a multiple of V_local (equal view groups)");
  MV_CHECK(group_index >= 0 && group_index < V_total / V_local, "group_index out of range");
  h->group_index = group_index;
  h->program = 0;
  MV_CUDA(cudaSetDevice(h->device));
  h->stream = (cudaStream_t)stream;
  h->v_total = V_total;
  h->kv_send = kv_send;
  h->kv_recv = kv_recv;
  h->kv_recv_bytes = (size_t)kv_recv_bytes;
  h->exchange = exchange;
  h->exchange_user = user;
  const std::vector<int> sv{V_local};
  Plan& p = h->plan_for(sv, H, W, true);
  cudaStream_t s = (cudaStream_t)stream;
  MV_CUDA(cudaMemcpyAsync(p.in_latents.p, latents, (size_t)V_local * h->cfg.in_channels * H * W * sizeof(float),
                          cudaMemcpyDeviceToDevice, s));
  MV_CUDA(cudaMemcpyAsync(p.in_t.p, timesteps, (size_t)V_local * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
  g_launch_count = 0;
  h->execute(p, s, nullptr);  // eager: the exchange callback re-enters the host
  p.launches = g_launch_count;
  MV_CUDA(cudaMemcpyAsync(out, p.out_eps.p, (size_t)V_local * h->cfg.out_channels * H * W * sizeof(float),
                          cudaMemcpyDeviceToDevice, s));
  h->last_launches = p.launches;
  MV_API_END
}

int mvldm_vae_decode(mvldm_handle h, void* stream, const float* latents, int n, int H, int W, float* image) {
  MV_API_BEGIN
  MV_CHECK(h && latents && image && n > 0 && H > 0 && W > 0, "bad argument");
  MV_CUDA(cudaSetDevice(h->device));
  h->forward((cudaStream_t)stream, latents, nullptr, std::vector<int>(1, n), H, W, image, 1);
  MV_API_END
}

int mvldm_vae_encode(mvldm_handle h, void* stream, const float* image, int n, int H, int W, float* moments) {
  MV_API_BEGIN
  MV_CHECK(h && image && moments && n > 0 && H > 0 && W > 0, "bad argument");
  MV_CUDA(cudaSetDevice(h->device));
  h->forward((cudaStream_t)stream, image, nullptr, std::vector<int>(1, n), H, W, moments, 2);
  MV_API_END
}

int mvldm_last_launch_count(mvldm_handle h) { return h ? h->last_launches : -1; }

int mvldm_set_profiling(mvldm_handle h, int enable) {
  MV_API_BEGIN
  MV_CHECK(h, "null handle");
  h->profiling = enable != 0;
  MV_API_END
}

const char* mvldm_profile_json(mvldm_handle h) {
  try {
    MV_CHECK(h, "null handle");
    return profile_report(h);
  } catch (const std::exception& e) {
    mvldm::g_last_error = e.what();
    return nullptr;
  }
}

int mvldm_enable_taps(mvldm_handle h, int enable) {
  MV_API_BEGIN
  MV_CHECK(h, "null handle");
  h->taps_enabled = enable != 0;
  MV_API_END
}

int mvldm_debug_tap(mvldm_handle h, void* stream, const char* name, float* out, int64_t* numel) {
  MV_API_BEGIN
  MV_CHECK(h && name && numel, "null argument");
  MV_CHECK(h->last_plan != nullptr, "no forward has run yet");
  auto it = h->last_plan->taps.find(name);
  MV_CHECK(it != h->last_plan->taps.end(), std::string("no such tap (enable taps and run a forward first): ") + name);
  const Act& a = it->second.a;
  *numel = a.tokens() * a.c;
  if (out) nhwc_to_nchw_f32((cudaStream_t)stream, a.p, a.n, a.h * a.w, a.c, out);
  MV_API_END
}

int mvldm_build_inputs(void* stream, const float* x_t, const float* ctx, const float* rays, int B, int v_c, int v_t,
                       int ray_views, int ray_off, int R, int hw, float* out) {
  MV_API_BEGIN
  MV_CHECK(x_t && rays && out && (ctx || v_c == 0), "null argument");
  MV_CHECK(ray_off + v_c + v_t <= ray_views, "ray views out of range");
  build_inputs((cudaStream_t)stream, x_t, ctx, rays, B, v_c, v_t, ray_views, ray_off, R, hw, out);
  MV_API_END
}

int mvldm_ddim_step(void* stream, const float* eps_c, const float* eps_u, float cfg_scale, int B, int v_c, int v_t,
                    int chw, const float* x_t, float sa, float s1a, float sp, float s1p, float* x_prev, float* eps_out) {
  MV_API_BEGIN
  MV_CHECK(eps_c && x_t && x_prev, "null argument");
  MV_CHECK(sa > 0.f, "sqrt(alpha_t) must be positive");
  ddim_step((cudaStream_t)stream, eps_c, eps_u, cfg_scale, B, v_c, v_t, chw, x_t, sa, s1a, sp, s1p, x_prev, eps_out);
  MV_API_END
}

int mvldm_raymap_encoded(void* stream, const float* extr, const float* intr, int n, int h, int w, int plucker,
                         int origin_octaves, int direction_octaves, int srt, float* out) {
  MV_API_BEGIN
  MV_CHECK(extr && intr && out, "null argument");
  MV_CHECK(origin_octaves >= 0 && origin_octaves <= 32 && direction_octaves >= 0 && direction_octaves <= 32, "octaves out of range");
  MV_CHECK(!srt || (origin_octaves > 0 && direction_octaves > 0), "the SRT ray encoder needs both octave counts > 0");
  raymap((cudaStream_t)stream, extr, intr, n, h, w, plucker != 0, out, origin_octaves, direction_octaves, srt != 0);
  MV_API_END
}

int mvldm_ddpm_step(void* stream, const float* eps_c, const float* eps_u, float cfg_scale, int B, int v_c, int v_t,
                    int chw, const float* x_t, const float* noise, float sa, float s1a, float c_x0, float c_xt, float sigma,
                    float clip, float* x_prev) {
  MV_API_BEGIN
  MV_CHECK(eps_c && x_t && x_prev, "null argument");
  MV_CHECK(sa > 0.f, "sqrt(alpha_t) must be positive");
  MV_CHECK(noise || sigma == 0.f, "a non-zero sigma needs the noise tensor");
  ddpm_step((cudaStream_t)stream, eps_c, eps_u, cfg_scale, B, v_c, v_t, chw, x_t, noise, sa, s1a, c_x0, c_xt, sigma, clip,
            x_prev);
  MV_API_END
}

int mvldm_raymap(void* stream, const float* extr, const float* intr, int n, int h, int w, int plucker, float* out) {
  MV_API_BEGIN
  MV_CHECK(extr && intr && out, "null argument");
  raymap((cudaStream_t)stream, extr, intr, n, h, w, plucker != 0, out);
  MV_API_END
}

int mvldm_op_gemm(void* stream, int impl, const mvldm_gemm_desc* d) {
  MV_API_BEGIN
  MV_CHECK(d, "null argument");
  if (impl == MVLDM_IMPL_TC) {
    static DevBuf scratch;  // op-level entry point (tests): grown on demand, never shrunk
    const size_t need = gemm_tc_workspace_bytes(*d);
    if (need > scratch.bytes) {
      MV_CUDA(cudaDeviceSynchronize());
      scratch.alloc(need);
      MV_CUDA(cudaMemset(scratch.p, 0, need));
    }
    gemm_tc((cudaStream_t)stream, *d, scratch.p, scratch.bytes);
  } else {
    gemm_simt((cudaStream_t)stream, *d);
  }
  MV_API_END
}

int mvldm_op_attention(void* stream, int impl, const void* qkv, void* out, int batches, int seq, int heads, int d,
                       int dpad) {
  MV_API_BEGIN
  MV_CHECK(qkv && out, "null argument");
  if (impl == MVLDM_IMPL_TC)
    attention_tc((cudaStream_t)stream, (const bf16*)qkv, (bf16*)out, batches, seq, heads, d, dpad);
  else
    attention_simt((cudaStream_t)stream, (const bf16*)qkv, (bf16*)out, batches, seq, heads, d, dpad);
  MV_API_END
}

int mvldm_debug_attn_trace(int64_t* out, int n) {
  MV_API_BEGIN
  MV_CHECK(out && n > 0 && n <= 8 * 512, "bad arguments");
  MV_CUDA(cudaDeviceSynchronize());
  attention_trace_read(reinterpret_cast<long long*>(out), n);
  MV_API_END
}

int mvldm_op_attention_kv(void* stream, const void* q, int ld_q, int q_col0, const void* kv, int ld_kv, int k_col0,
                          int v_col0, void* out, int batches, int seq_q, int seq_kv, int heads, int d, int dpad,
                          float* stats) {
  MV_API_BEGIN
  MV_CHECK(q && kv && out, "null argument");
  attention_tc_kv((cudaStream_t)stream, (const bf16*)q, ld_q, q_col0, (const bf16*)kv, ld_kv, k_col0, v_col0, (bf16*)out,
                  batches, seq_q, seq_kv, heads, d, dpad, stats);
  MV_API_END
}

int mvldm_op_attention_merge(void* stream, int nparts, const void* const* parts, const float* const* stats, int64_t rows,
                             int heads, int dpad, void* out) {
  MV_API_BEGIN
  MV_CHECK(parts && stats && out && nparts >= 1 && nparts <= 3, "bad arguments");
  for (int i = 0; i < nparts; ++i) MV_CHECK(parts[i] && stats[i], "null part");
  attention_merge((cudaStream_t)stream, nparts, reinterpret_cast<const bf16* const*>(parts), stats, rows, heads, dpad,
                  (bf16*)out);
  MV_API_END
}

int mvldm_op_groupnorm(void* stream, const void* x0, int c0, const void* x1, int c1, int n_img, int hw, int groups,
                       float eps, const float* gamma, const float* beta, int silu, void* out, float* scratch) {
  MV_API_BEGIN
  MV_CHECK(x0 && gamma && beta && out && scratch, "null argument");
  groupnorm_init();
  groupnorm((cudaStream_t)stream, (const bf16*)x0, c0, (const bf16*)x1, c1, n_img, hw, groups, eps, gamma, beta,
            silu != 0, (bf16*)out, scratch);
  MV_API_END
}

int mvldm_op_layernorm(void* stream, const void* x, int rows, int c, float eps, const float* gamma, const float* beta,
                       void* out) {
  MV_API_BEGIN
  MV_CHECK(x && gamma && beta && out, "null argument");
  layernorm((cudaStream_t)stream, (const bf16*)x, rows, c, eps, gamma, beta, (bf16*)out);
  MV_API_END
}

}  // extern "C"
