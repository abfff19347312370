// Op descriptors of the fused sequence kernel (seq.cu): one persistent launch executes a list of ops - implicit-GEMM
// convs / linears on tcgen05, GroupNorm(+SiLU), LayerNorm, the small elementwise helpers - with a grid-wide barrier
// between consecutive ops instead of a kernel boundary.  The host (model.cu) records the denoiser forward as a few
// such lists, cut only at the attention launches.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace mvldm {

constexpr int SEQ_KC = 2;         // 64-channel K chunks one pipeline step (one TMA box per operand) carries
constexpr int SEQ_THREADS = 256;  // warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue, all 8 for elementwise ops
constexpr int SEQ_MAX_SPLITS = 32;

enum { SEQ_GEMM = 0, SEQ_GN = 1, SEQ_LN = 2, SEQ_UPSAMPLE = 3, SEQ_IM2COL = 4, SEQ_SINUSOID = 5, SEQ_SPLITK_REDUCE = 6 };

struct SeqSeg {  // one K-segment of the implicit A operand (mvldm_aseg, compacted)
  int16_t ncblk, ntaps, stride, spt, kchunk0, pad_;
  int8_t dh[9], dw[9];
  int16_t cblk[9];  // tap channel offset / 64
  int16_t pad2_;
};

struct SeqGemm {
  SeqSeg seg[MVLDM_MAX_SEGS];
  int nseg, M, N, num_steps;
  int mt, nt, splits, steps_per_split;  // work items = mt * nt * splits, m fastest
  int hw, ow;                           // output pixels per image / row width
  int bn, stages;                       // N tile (multiple of 32, <= 256) and smem ring depth for it
  int solo, pad_;                       // solo: at most one work item per CTA (all 8 warps then run the epilogue)
  int mode, ldo, n_valid, rowvec_ld, res_ld;
  int stage_out;  // one work item per CTA and a bf16 output: the tile leaves through shared memory as whole 128-byte lines
  float* partial;   // split-K fp32 partial tiles [splits][M][N], or NULL
  int* counters;    // fused split-K reduction: [2][mt*nt] arrive / done counters (zero between ops), or NULL
  const float* bias;
  const float* rowvec;
  const bf16* residual;
  void* out;
};

struct SeqGN {
  const bf16* x0;
  const bf16* x1;
  const float* gamma;
  const float* beta;
  bf16* out;
  float* partial;  // [items][4][2] (mean, M2) per pixel chunk when ps > 1
  int c0, c1, n_img, hw;
  int cgn, cb;     // channels per group; channels per item block = lcm(cgn, 8): whole groups, whole 16-byte vectors
  int ps, px;      // pixel chunks per image and pixels per chunk (ps > 1: statistics meet across CTAs at a grid barrier)
  int silu, gw;    // gw: warps that own one item (1, 2, 4 or 8; 8 / gw items are in flight per CTA)
  int n_items, pad_;
  float eps;
};

struct SeqLN {
  const bf16* x;
  const float* gamma;
  const float* beta;
  bf16* out;
  int rows, c;
  float eps;
};

struct SeqEW {  // upsample / im2col / sinusoid / split-K reduce
  const void* src;
  const void* src2;   // split-K reduce: residual
  void* dst;
  const float* bias;
  const float* rowvec;
  int n_img, h, w, c;  // upsample: source size; im2col: cin = c, kpad = aux; sinusoid: n = n_img, dim = c
  int aux, aux2, aux3, aux4;
};

struct SeqPrefetch {  // weight tiles of the NEXT GEMM op, pulled into L2 by the (by then idle) producer lane of this GEMM op
  const CUtensorMap* map;  // that op's 3-D weight map, box [64 k, bn rows, SEQ_KC chunks]
  int bn, nt, nchunks, pad_;  // its N tile, N tiles and 64-wide K chunks
};

struct alignas(16) SeqOpC {  // the part every thread reads: staged in shared memory at each op boundary
  int type, index;
  int next_nseg;  // K-segments of the next op of this launch if it is a GEMM (its tensor maps are prefetched one op ahead), else 0
  int pad_;
  union {
    SeqGemm g;
    SeqGN gn;
    SeqLN ln;
    SeqEW ew;
  };
  SeqPrefetch pf;
};
constexpr int SEQ_OPC_BYTES = 512;
static_assert(sizeof(SeqOpC) <= SEQ_OPC_BYTES, "compact op descriptor must fit its shared-memory slot");

struct alignas(128) SeqOp {
  SeqOpC c;
  uint8_t pad_[SEQ_OPC_BYTES - sizeof(SeqOpC)];
  CUtensorMap tmA[MVLDM_MAX_SEGS][SEQ_KC];  // 5-D (64 ch, w, h, image, chunk), box carrying 1..KC chunks
  CUtensorMap tmB[SEQ_KC];                   // 3-D (64 k, n, chunk)
  CUtensorMap tmPf;                          // copy of the next GEMM op's tmB[SEQ_KC - 1]: weight prefetch of a single-op launch
};

// ---- host side ------------------------------------------------------------------------------------------
// Fill `op` for one GEMM (tile / split-K choice, tensor maps).  `workspace` = split-K scratch (counters + fp32 partials) shared
// by all ops of a forward, or NULL.  Returns true when a separate SEQ_SPLITK_REDUCE op must follow (`reduce` filled).
bool seq_plan_gemm(const mvldm_gemm_desc& d, void* workspace, size_t workspace_bytes, SeqOp& op, SeqOp& reduce);
size_t seq_gemm_workspace_bytes(const mvldm_gemm_desc& d);
void seq_plan_groupnorm(const bf16* x0, int c0, const bf16* x1, int c1, int n_img, int hw, int groups, float eps,
                        const float* gamma, const float* beta, bool silu, bf16* out, float* scratch, int grid, SeqOp& op);
size_t seq_groupnorm_scratch_floats(int n_img, int c, int groups);
void seq_plan_layernorm(const bf16* x, int rows, int c, float eps, const float* gamma, const float* beta, bf16* out, SeqOp& op);
void seq_plan_upsample(const bf16* x, int n_img, int h, int w, int c, bf16* out, SeqOp& op);
void seq_plan_im2col(const float* latents, int n_img, int cin, int h, int w, int kpad, bf16* out, SeqOp& op);
void seq_plan_sinusoid(const int64_t* t, int n, int dim, bf16* out, SeqOp& op);
// after all ops of a forward are planned: every GEMM op gets the L2 prefetch plan of the next GEMM op's weights (also across
// launch boundaries: the attention kernel in between leaves them in L2)
void seq_link_prefetch(SeqOp* ops, int n, const SeqOp* dev_ops);
// per launch (ops[0..n) run in one sequence launch): tensor-map prefetch hints one op ahead
void seq_link_launch(SeqOp* ops, int n);

void seq_configure();  // kernel attributes, once per device (call outside stream capture)
int seq_grid();  // CTAs of every sequence launch (= SM count; all co-resident)
// `sync`: two zero-initialised device words owned by the caller (re-armed by the kernel); `timing`: NULL or n_ops + 1 device
// int64 (globaltimer ns at kernel start and after every op's barrier, written by CTA 0)
void seq_launch(cudaStream_t s, const SeqOp* dev_ops, int n_ops, unsigned* sync, long long* timing);
// one-off launch of host-built ops (op-level C ABI entry points / tests): uploads to a cached device buffer
void seq_run_host_ops(cudaStream_t s, const SeqOp* host_ops, int n_ops);
// ONE GEMM op as its own launch: the descriptor and tensor maps travel as kernel parameters (no grid barrier, no global
// descriptor fetch); min(work items, SMs) CTAs
void seq_launch_gemm(cudaStream_t s, const SeqOp& op);
void seq_debug_empty_ops(cudaStream_t s, int n_ops);
extern long long* g_seq_trace;  // debug: [barrier][cta][4] timeline buffer for the next launches, or NULL

}  // namespace mvldm
