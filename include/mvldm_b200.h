/*
 * mvldm_b200 — C ABI of the B200-native MV-LDM denoising hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every entry point returns 0 on success and a non-zero
 * code on failure; mvldm_last_error() then returns a thread-local message (reference error behaviour is
 * Python exceptions/asserts, e.g. src/model/diffusion_wrapper.py:149,673 — the Python host raises
 * RuntimeError from this string).  All device work is enqueued asynchronously on the caller's
 * cudaStream_t (passed as void*); nothing here synchronises the device, so every call is CUDA-graph
 * capturable.  There is no CPU fallback: a missing GPU / unsupported shape is an error.
 *
 * Each function names the reference interface it replaces (paths relative to the reference repo).
 */
#ifndef MVLDM_B200_H_
#define MVLDM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MVLDM_MAX_LEVELS 4
#define MVLDM_MAX_SEGS 3

typedef struct mvldm_handle_s* mvldm_handle;

/* dtype codes for mvldm_set_weight */
enum { MVLDM_F32 = 0, MVLDM_BF16 = 1, MVLDM_F16 = 2 };
/* kernel family: tcgen05/TMA kernels (product) or the plain CUDA-core kernels kept as an on-device
 * cross-check for the tests (never selected implicitly) */
enum { MVLDM_IMPL_TC = 0, MVLDM_IMPL_SIMT = 1, MVLDM_IMPL_TC_GEMM_SIMT_ATTN = 2 };
enum { MVLDM_MV_SPATIAL_TRANSFORMER_3D = 0, MVLDM_MV_STANDARD = 1 };
enum { MVLDM_MODEL_DENOISER = 0, MVLDM_MODEL_VAE = 1 };

/* Replaces: MultiViewUNetCfg + UNet2DModelCfg + SpatialTransformer3DCfg
 * (src/model/denoiser/mvunet.py:22-40, src/model/denoiser/mvdream/attention.py:23-32) and the
 * UNet2DConditionModel constructor arguments at mvunet.py:54-63. */
typedef struct {
  int32_t in_channels;                          /* 4 latent + 1 mask + ray channels (diffusion_wrapper.py:98-129) */
  int32_t out_channels;
  int32_t num_levels;                           /* len(block_out_channels) */
  int32_t block_out_channels[MVLDM_MAX_LEVELS]; /* config/model/denoiser/mv_unet.yaml:10 */
  int32_t layers_per_block;                     /* diffusers default 2 */
  int32_t norm_groups;                          /* 32 */
  int32_t num_heads;                            /* multi_view_attention.num_heads */
  int32_t max_attn_res;                         /* mvunet.py:137,190: multi-view block only if h,w <= 32 */
  int32_t impl;                                 /* MVLDM_IMPL_* */
  int32_t use_cuda_graph;                       /* capture one graph per (B,V,h,w) and replay it */
  /* Variant B = `pretrained_from` set (mvunet.py:51-59): the SD-2.1 UNet2DConditionModel topology.  Down blocks
   * 0..L-2 and the mid block carry a per-view Transformer2DModel after each resnet (mvunet.py:118-131,150-160) whose
   * cross-attention sees one all-zero token; the up blocks' attentions exist in the state dict but are never run
   * (mvunet.py:166-186 calls only the resnets), so their keys are accepted and dropped. */
  int32_t variant;                              /* 0 = A (DownBlock2D/UpBlock2D), 1 = B (SD-2.1 topology) */
  int32_t t2d_heads[MVLDM_MAX_LEVELS];          /* SD-2.1 attention_head_dim [5,10,20,20] (= heads; head dim 64) */
  int32_t cross_attention_dim;                  /* 1024 */
  /* multi_view_attention.name (src/model/denoiser/attention.py:8-27): which block sits at the 9 multi-view positions.
   * 1 = StandardTransformer (src/model/denoiser/standard/transformer.py:45-136, the reference's default
   * config/model/denoiser/multi_view_attention/standard_attention.yaml): per layer  x += to_out(softmax(q k^T / sqrt(d)) v)
   * over ALL (view, pixel) tokens of a scene on LayerNorm(x), then x += W2 gelu(W1 LayerNorm(x)); no GroupNorm, no
   * proj_in / proj_out.  Supported: d_dot = d_in / num_heads, downscale = 1, pos_enc = false. */
  int32_t mv_block;                             /* MVLDM_MV_SPATIAL_TRANSFORMER_3D or MVLDM_MV_STANDARD */
  int32_t mv_num_layers;                        /* CrossAttentionCfg.num_layers (standard only; >= 1) */
  int32_t mv_d_mlp;                             /* CrossAttentionCfg.d_mlp, or 0 to use the multiplier */
  int32_t mv_d_mlp_multiplier;                  /* hidden width = d_in * multiplier when mv_d_mlp == 0 */
  /* MVLDM_MODEL_VAE: the handle is the first-stage autoencoder instead of the denoiser - diffusers AutoencoderKL as built by
   * get_autoencoder (src/model/autoencoder/__init__.py:36-43; SD-2.1 "vae" subfolder: block_out_channels [128,256,512,512],
   * layers_per_block 2, 32 groups, latent_channels 4).  Uses in_channels / out_channels (image channels), num_levels,
   * block_out_channels, layers_per_block, norm_groups, latent_channels; state-dict keys are diffusers' (encoder.*,
   * decoder.*, quant_conv.*, post_quant_conv.*).  Entry points: mvldm_vae_encode / mvldm_vae_decode. */
  int32_t model;                                /* MVLDM_MODEL_DENOISER or MVLDM_MODEL_VAE */
  int32_t latent_channels;                      /* 4 */
} mvldm_config;

const char* mvldm_last_error(void);
int mvldm_version(void);

/* Replaces MultiViewUNet.__init__ (mvunet.py:43-88): builds the layer table, no weights yet. */
int mvldm_create(const mvldm_config* cfg, int device, mvldm_handle* out);
int mvldm_destroy(mvldm_handle h);

/* Weight ingestion.  Keys are the reference module's state_dict keys without the Lightning
 * "denoiser." prefix, e.g. "unet.down_blocks.0.resnets.1.conv1.weight",
 * "cross_attn_blocks_mid.0.transformer_blocks.0.attn1.to_q.weight".  `ptr` is a DEVICE pointer to a
 * contiguous tensor of `dtype`; it is copied, the caller keeps ownership.
 * Replaces nn.Module.load_state_dict for the denoiser (scripts/generate_mvldm.py:66). */
int mvldm_num_weights(mvldm_handle h);
const char* mvldm_weight_name(mvldm_handle h, int index);
int mvldm_weight_shape(mvldm_handle h, int index, int64_t shape[4], int* ndim);
int mvldm_set_weight(mvldm_handle h, const char* key, const void* ptr, const int64_t* shape, int ndim,
                     int dtype, void* stream);
/* Packs every layer into its kernel layout (bf16, tap-major conv filters, fused QKV with padded heads,
 * conv2+shortcut K-concatenation, GEGLU column interleave).  Fails if a key is missing. */
int mvldm_finalize_weights(mvldm_handle h, void* stream);

/* Bytes of device workspace the handle holds for a (B scenes, V views, h x w latent) call. */
int64_t mvldm_workspace_bytes(mvldm_handle h, int B, int V, int H, int W);

/* SUPPORTED SHAPE ENVELOPE (anything outside raises through mvldm_last_error; there is no fallback kernel):
 *   latent H, W     multiples of 2^(num_levels-1) (8 for the 4-level UNet).  The implicit-GEMM M tile is a TMA box of
 *                   (columns x rows x images) covering <= 128 rows of the 128-row MMA (csrc/gemm_tc.cu tile_geometry()):
 *                   widths that divide 128 with h*w | 128 or 128 | h*w tile densely (32x32, 16x16, 64x64, 32x16, 8x8 ...:
 *                   the fast path); any other width <= 128 (24, 48, 40, 12, 6, 5, 3 ...) uses the box with the fewest
 *                   tiles whose row count is a multiple of 8, e.g. 24 columns x 5 rows = 120 rows; widths > 128 are split
 *                   into equal column blocks (256 = 2 x 128, 192 = 2 x 96).  Tested: 32x32, 16x16, 64x64, 32x16, 8x8,
 *                   24x24, 48x40, 40x24 latents and 256- / 192-wide feature maps at the op level.  Refused: widths with
 *                   no equal split into <= 128-pixel blocks whose box has a multiple-of-8 row count (e.g. 129).
 *   channels        block_out_channels multiples of 64 (TMA K-block), divisible by norm_groups (<= 64 groups) and by
 *                   num_heads; concatenated resnet inputs <= 2560 channels; head dim <= 192 (padded to 64/128/192).
 *   views / scenes  any B >= 1, V >= 1 (32-bit indexing: B*V*H*W*C < 2^31 per tensor); multi-view blocks run only at
 *                   levels with h_l, w_l <= max_attn_res (mvunet.py:137,190).
 *   in/out channels any in_channels (conv_in runs on an explicit im2col operand), out_channels <= 32.
 *   dtypes          inputs / outputs fp32, timesteps int64, weights fp32/bf16/fp16 at ingestion; arithmetic bf16
 *                   operands with fp32 accumulation and fp32 normalisation statistics.
 *
 * Replaces Denoiser.forward / MultiViewUNet.forward (src/model/denoiser/denoiser.py:22-29,
 * mvunet.py:90-208).  latents: device fp32 [B,V,in_channels,H,W] contiguous; timesteps: device int64
 * [B*V] (the [B] form of mvunet.py:102-105 is expanded by the host); out: device fp32
 * [B,V,out_channels,H,W]. */
int mvldm_forward(mvldm_handle h, void* stream, const float* latents, const int64_t* timesteps, int B, int V,
                  int H, int W, float* out);
/* Several scenes with DIFFERENT view counts in one pass: `latents` is [sum(views_per_scene), C_in, H, W] with the
 * views of a scene contiguous, `timesteps` one int64 per view.  Only the joint attention looks across views, and it
 * runs per scene; everything else is per view.  This is how one DDIM step with classifier-free guidance becomes ONE
 * forward: the conditional pass (v_c + v_t views, diffusion_wrapper.py:429-435) and the unconditional pass (v_t views,
 * :437-441) of every scene go through the network together. */
int mvldm_forward_scenes(mvldm_handle h, void* stream, const float* latents, const int64_t* timesteps, int num_scenes,
                         const int32_t* views_per_scene, int H, int W, float* out);

/* First-stage autoencoder (handle created with cfg.model = MVLDM_MODEL_VAE).
 * mvldm_vae_decode replaces `self.autoencoder.decode(latents).sample` (src/model/diffusion_wrapper.py:293-295): latents device
 * fp32 [n, latent_channels, H, W] (already divided by the 0.18215 scaling, as the caller does at :292) -> image fp32
 * [n, out_channels, H*f, W*f], f = 2^(num_levels-1), values nominally in [-1, 1] (the caller maps to [0, 1], :298).
 * mvldm_vae_encode replaces `self.autoencoder.encode(inputs).latent_dist` (:283): image fp32 [n, in_channels, H, W] in
 * [-1, 1] -> moments fp32 [n, 2*latent_channels, H/f, W/f] = (mean | logvar) of the diagonal Gaussian; the caller draws the
 * sample (mean + exp(0.5 * clamp(logvar, -30, 20)) * noise) so that the host RNG stays the reference's.
 * H*W/f^2 (tokens per image in the mid-block attention) must be a multiple of 64 and <= 4096. */
int mvldm_vae_decode(mvldm_handle h, void* stream, const float* latents, int n, int H, int W, float* image);
int mvldm_vae_encode(mvldm_handle h, void* stream, const float* image, int n, int H, int W, float* moments);

/* View-group sharding: a scene whose views are split over several GPUs.  This rank holds the
 * `group_index`-th group of V_local contiguous views of ONE scene (B = 1) out of V_total; every op is per view except
 * the joint multi-view attention, which needs all views' K and V.  At each of the 9 multi-view blocks the library
 *   1. packs the local K|V (bf16 [V_local*h*w, 2*heads*dpad]) into `kv_send`,
 *   2. calls exchange(..., MVLDM_EXCHANGE_BEGIN): the host STARTS the all-gather (NCCL over NVLink) - on a stream of its
 *      own, ordered after the work already enqueued on `stream` - and returns without waiting,
 *   3. runs the attention of this rank's queries against its OWN keys (from kv_send) while the slabs are on the wire,
 *   4. calls exchange(..., MVLDM_EXCHANGE_END): the host makes `stream` wait until `kv_recv` holds every rank's slab in
 *      view order (bf16 [V_total*h*w, 2*heads*dpad]),
 *   5. runs the attention against the slabs in front of and behind its own and merges the (up to) three partial
 *      softmaxes in that fixed order (own, before, after): deterministic, and equal to the one-pass softmax up to
 *      the rounding of the merge.
 * A host that does not overlap runs the whole exchange on `stream` inside BEGIN and returns MVLDM_EXCHANGE_DONE instead of 0:
 * the library then makes ONE pass over all slabs in view order (no partial softmaxes, no END call).  Measured on NVLink 5
 * (1 scene x 64 views, profiles/r02_final_bench_n{2,4,8}.json) the exchange is a few percent of the forward and the
 * three-way split costs about what it hides, so the one-pass mode is the host's default.  Not graph-captured
 * (the callback re-enters the host).  kv_send must hold V_local*h*w*2*heads*64*2 bytes at the finest level, kv_recv
 * V_total/V_local times that. */
enum { MVLDM_EXCHANGE_BEGIN = 0, MVLDM_EXCHANGE_END = 1 };
enum { MVLDM_EXCHANGE_DONE = 2 };  /* return value of the BEGIN call: exchange already complete on `stream` */
typedef int (*mvldm_kv_exchange_fn)(void* user, const void* kv_send, void* kv_recv, int64_t bytes_per_rank, void* stream,
                                    int phase);
int mvldm_forward_sharded(mvldm_handle h, void* stream, const float* latents, const int64_t* timesteps, int V_local,
                          int V_total, int group_index, int H, int W, float* out, void* kv_send, void* kv_recv,
                          int64_t kv_recv_bytes, mvldm_kv_exchange_fn exchange, void* user);

/* Number of kernels the last mvldm_forward on this handle launched (graph replays count their nodes). */
int mvldm_last_launch_count(mvldm_handle h);

/* Per-op device timing for bench.py's roofline: while enabled, mvldm_forward runs eagerly (no graph) with a
 * CUDA-event pair around every op on the launching stream; mvldm_profile_json synchronises and returns
 * {"categories": {name: {launches, us, gflop, mbytes}}, "ops": [...]} for the last forward (string owned by
 * the handle, valid until the next call). */
int mvldm_set_profiling(mvldm_handle h, int enable);
const char* mvldm_profile_json(mvldm_handle h);

/* Copy an intermediate activation of the last (non-graph) forward into `out` as fp32 NCHW
 * [B*V, C, h, w] for per-layer parity tests; names follow oracle taps ("down0.res1", "mid.mv", ...).
 * Returns the element count through *numel (out may be NULL to query). */
int mvldm_debug_tap(mvldm_handle h, void* stream, const char* name, float* out, int64_t* numel);
int mvldm_enable_taps(mvldm_handle h, int enable);

/* Replaces the three torch.cat of DiffusionWrapper.step (diffusion_wrapper.py:429-432,438):
 * inputs[b, v] = [latent(4) | mask(1) | rays(R)], context views first.  All fp32 device pointers:
 * x_t [B,v_t,4,hw], context_latents [B,v_c,4,hw] (may be NULL iff v_c == 0), rays [B,v_c_total+v_t,R,hw]
 * where the first `ray_view_offset` views of `rays` are skipped.  Mask = 0 for context, 1 for target
 * (diffusion_wrapper.py:476-477).  out [B, v_c+v_t, 5+R, hw]. */
int mvldm_build_inputs(void* stream, const float* x_t, const float* context_latents, const float* rays,
                       int B, int v_c, int v_t, int ray_views, int ray_view_offset, int R, int hw, float* out);

/* Replaces the CFG compose (diffusion_wrapper.py:444/447) fused with DDIMScheduler.step
 * (diffusers, eta=0, epsilon prediction, no clipping; called at diffusion_wrapper.py:451):
 *   eps = eps_u + s*(eps_c[:, v_c:] - eps_u)   (eps_u == NULL: eps = eps_c[:, v_c:])
 *   x0  = (x_t - sqrt(1-a_t) eps) / sqrt(a_t);   x_prev = sqrt(a_prev) x0 + sqrt(1-a_prev) eps
 * eps_c [B, v_c+v_t, chw], eps_u [B, v_t, chw] or NULL, x_t/x_prev [B, v_t, chw], all fp32. */
int mvldm_ddim_step(void* stream, const float* eps_c, const float* eps_u, float cfg_scale, int B, int v_c,
                    int v_t, int chw, const float* x_t, float sqrt_a_t, float sqrt_1m_a_t, float sqrt_a_prev,
                    float sqrt_1m_a_prev, float* x_prev, float* eps_out /* may be NULL */);

/* Replaces diffusers DDPMScheduler.step (the "ddpm" entry of SCHEDULER, src/model/scheduler/__init__.py:19-22,
 * config/model/scheduler/ddpm.yaml) fused with the CFG compose: x0 = (x_t - s1a*eps)/sa, clamped to +-clip when clip > 0
 * (clip_sample / clip_sample_range), x_prev = c_x0*x0 + c_xt*x_t + sigma*noise.  The host computes the scalars
 * (c_x0 = sqrt(a_prev)*beta_t/(1-a_t), c_xt = sqrt(alpha_t)*(1-a_prev)/(1-a_t), sigma = sqrt(variance), 0 at t = 0) and
 * draws `noise` (fp32, x_t's shape, NULL when sigma == 0) on the caller's generator. */
int mvldm_ddpm_step(void* stream, const float* eps_c, const float* eps_u, float cfg_scale, int B, int v_c, int v_t,
                    int chw, const float* x_t, const float* noise, float sa, float s1a, float c_x0, float c_xt, float sigma,
                    float clip, float* x_prev);

/* Replaces DiffusionWrapper.ray_encode / generate_image_rays (diffusion_wrapper.py:169-190,301-322) with
 * get_world_rays / sample_image_grid (src/geometry/projection.py:91-138) for use_ray_encoding=false,
 * srt_ray_encoding=false: extr [n,4,4] cam-to-world, intr [n,3,3] normalised (fp32, device);
 * out [n,6,h,w] = origin (or origin x direction when plucker) then unit direction. */
int mvldm_raymap(void* stream, const float* extr, const float* intr, int n_views, int h, int w, int plucker,
                 float* out);
/* use_ray_encoding=true (config/main.yaml:28-33; diffusion_wrapper.py:115-126,317-320): the origin and the direction triple
 * are each replaced by their PositionalEncoding (src/model/encodings/positional_encoding.py:28-49) when the octave count is
 * > 0: out [n, C, h, w], C = (6*origin_octaves or 3) + (6*direction_octaves or 3), channel (d f p) = sin(2 pi 2^f x_d + p pi/2).
 * srt != 0: srt_ray_encoding=true (diffusion_wrapper.py:104-113,311-315) - RayEncoder (src/model/srt/layers.py:9-58): per triple
 * [sin(pi 2^f x_d) over (d f) | cos(pi 2^f x_d) over (d f)], origins then directions; both octave counts must be > 0. */
int mvldm_raymap_encoded(void* stream, const float* extr, const float* intr, int n_views, int h, int w, int plucker,
                         int origin_octaves, int direction_octaves, int srt, float* out);

/* ---------------------------------------------------------------------------------------------
 * Op-level entry points: the kernels behind mvldm_forward, exposed so tests can check each against
 * the oracle.  bf16 tensors are raw uint16 device buffers in NHWC ([images, h, w, channels]).
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const void* ptr;      /* bf16 NHWC source [n_img, sh, sw, ctot] */
  int32_t c;            /* channels consumed per tap (multiple of 64 for the tcgen05 path) */
  int32_t ctot;         /* channel count (row pitch) of the source tensor */
  int32_t sh, sw;       /* source spatial size */
  int32_t stride;       /* 1 or 2 */
  int32_t ntaps;        /* 1 or 9 */
  int8_t dh[9], dw[9];  /* tap t reads source pixel (oh*stride + dh[t], ow*stride + dw[t]), zero outside */
  int32_t coff[9];      /* ... at channel coff[t] + c */
} mvldm_aseg;

typedef struct {
  /* A (implicit): K = sum over segments of ntaps*c, ordered segment-major, tap, channel */
  int32_t nseg;
  mvldm_aseg seg[MVLDM_MAX_SEGS];
  int32_t n_img, oh, ow;        /* M = n_img*oh*ow output pixels */
  /* B: bf16 [N, K] row-major (K contiguous) */
  const void* w;
  int32_t n, k;
  /* epilogue: acc + bias[n] + rowvec[img, n] + residual[m, n] */
  const float* bias;            /* [N] or NULL */
  const float* rowvec;          /* [n_img, rowvec_ld] or NULL */
  int32_t rowvec_ld;
  const void* residual;         /* bf16 [M, res_ld] or NULL */
  int32_t res_ld;
  int32_t mode;                 /* 0 bf16 [M,ldo]; 1 GEGLU (8-col interleave: columns [16j,16j+8) values, [16j+8,16j+16) their gates) -> bf16 [M, N/2]; 2 fp32 NCHW, n_valid channels;
                                   3 like 0 with SiLU on the result; 4 fp32 [M,ldo]; 5 like 0 with exact (erf) GELU on the result */
  void* out;
  int32_t ldo;
  int32_t n_valid;
} mvldm_gemm_desc;

int mvldm_op_gemm(void* stream, int impl, const mvldm_gemm_desc* d);

/* Self-attention over packed, head-padded q|k|v: qkv bf16 [batches*seq, 3*heads*dpad] (q block, k block,
 * v block; head h at columns h*dpad, first d of dpad valid; pad columns zero EXCEPT column d of every V head, which
 * must hold 1.0: the tcgen05 kernel reads the softmax row sum from P.V through it; dpad > d), out bf16
 * [batches*seq, heads*dpad] (pad columns zero).  softmax(q k^T d^-1/2) v with fp32 scores (mvdream/attention.py:174-205). */
int mvldm_op_attention(void* stream, int impl, const void* qkv, void* out, int batches, int seq, int heads,
                       int d, int dpad);
/* Same (tcgen05 path only) with queries and keys/values in different buffers and of different lengths:
 * q bf16 [batches*seq_q, ld_q] (head h at column q_col0 + h*dpad), kv bf16 [batches*seq_kv, ld_kv] (K heads at
 * k_col0 + h*dpad, V heads at v_col0 + h*dpad, ones column as above).
 * `stats` (NULL or fp32 [batches*seq_q, heads, 2]): also report every row's softmax state (reference max in the log2
 * domain, row sum against it), so that launches over DISJOINT key ranges can be combined by mvldm_op_attention_merge. */
int mvldm_op_attention_kv(void* stream, const void* q, int ld_q, int q_col0, const void* kv, int ld_kv, int k_col0,
                          int v_col0, void* out, int batches, int seq_q, int seq_kv, int heads, int d, int dpad,
                          float* stats);
/* out = softmax over the union of the key ranges of 1..3 partial launches: parts[i] bf16 [rows, heads*dpad] with
 * stats[i] from mvldm_op_attention_kv; combined in argument order (bit-stable). */
int mvldm_op_attention_merge(void* stream, int nparts, const void* const* parts, const float* const* stats, int64_t rows,
                             int heads, int dpad, void* out);

/* Debug: clock64() stamps of CTA (0,0) of the last attention launch made with MVLDM_ATTN_TRACE set in the
 * environment; out[slot*512 + tile], slots 0-2 softmax thread (wait S, got S, P handed over), 3-5 MMA thread
 * (got P, PV issued, next QK issued). */
int mvldm_debug_attn_trace(int64_t* out, int n);
/* GroupNorm (+SiLU) over NHWC bf16, optionally over the channel concat of two sources
 * (torch.cat at mvunet.py:176 + ResnetBlock2D.norm1): out bf16 [n_img, hw, c0+c1]. */
int mvldm_op_groupnorm(void* stream, const void* x0, int c0, const void* x1, int c1, int n_img, int hw,
                       int groups, float eps, const float* gamma, const float* beta, int silu, void* out,
                       float* scratch /* >= n_img*groups*2*64 floats */);
int mvldm_op_layernorm(void* stream, const void* x, int rows, int c, float eps, const float* gamma,
                       const float* beta, void* out);
#ifdef __cplusplus
}
#endif
#endif /* MVLDM_B200_H_ */
