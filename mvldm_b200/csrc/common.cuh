// Shared declarations for the mvldm_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/mvldm_b200.h"

typedef __nv_bfloat16 bf16;

namespace mvldm {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] void fail(const char* file, int line, const std::string& msg);

#define MV_CHECK(cond, msg)                                        \
  do {                                                             \
    if (!(cond)) ::mvldm::fail(__FILE__, __LINE__, std::string(msg)); \
  } while (0)

#define MV_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      ::mvldm::fail(__FILE__, __LINE__, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// launch counter (mvldm_last_launch_count): every kernel launch in this library goes through MV_LAUNCHED()
extern thread_local int g_launch_count;
#define MV_LAUNCHED()                 \
  do {                                \
    ++::mvldm::g_launch_count;        \
    MV_CUDA(cudaPeekAtLastError());   \
  } while (0)

// Function attributes (max dynamic shared memory, ...) belong to a device: a process that drives several GPUs must set
// them once per device, not once per process.  `flags` is a zero-initialised static array owned by the call site.
constexpr int kMaxDevices = 64;
inline bool first_use_on_device(bool (&flags)[kMaxDevices]) {
  int dev = 0;
  MV_CUDA(cudaGetDevice(&dev));
  MV_CHECK(dev >= 0 && dev < kMaxDevices, "device index out of range");
  if (flags[dev]) return false;
  flags[dev] = true;
  return true;
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every forward-path kernel is launched with programmaticStreamSerialization: it may be scheduled while its
// predecessor is still draining, runs its prologue (barrier init, TMEM alloc, descriptor prefetch), then
// blocks in pdl_wait() until the predecessor's memory is visible.  No kernel touches global memory before
// pdl_wait(), so RAW and WAR hazards through the reused activation arena are preserved.
extern bool g_use_pdl;  // MVLDM_PDL=1 enables; default is plain stream order
#ifdef __CUDACC__
#ifdef MVLDM_ENABLE_PDL  // compile-time opt-in (measured slower, profiles/r01_notes.md); otherwise the hooks vanish
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#else
__device__ __forceinline__ void pdl_wait() {}
__device__ __forceinline__ void pdl_launch_dependents() {}
#endif

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
#ifdef MVLDM_ENABLE_PDL
  cfg.numAttrs = g_use_pdl ? 1 : 0;
#else
  cfg.numAttrs = 0;
#endif
  MV_CUDA(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...));
  MV_LAUNCHED();
}
#endif

// ---- kernels (host launchers) -----------------------------------------------------------------
int sm_count();  // SMs of the current device
// gemm_simt.cu / gemm_tc.cu
void gemm_simt(cudaStream_t s, const mvldm_gemm_desc& d);
// workspace: fp32 split-K scratch (gemm_tc_workspace_bytes(d) bytes, zero-initialised) or NULL/0 to force a single pass
size_t gemm_tc_workspace_bytes(const mvldm_gemm_desc& d);
void gemm_tc(cudaStream_t s, const mvldm_gemm_desc& d, void* workspace, size_t workspace_bytes);
// attn_simt.cu / attn_tc.cu
void attention_simt(cudaStream_t s, const bf16* qkv, bf16* out, int batches, int seq, int heads, int d, int dpad);
void attention_trace_read(long long* host, int n);  // debug timeline of CTA (0,0), see MVLDM_ATTN_TRACE
void attention_tc(cudaStream_t s, const bf16* qkv, bf16* out, int batches, int seq, int heads, int d, int dpad);
// queries and keys/values from different buffers / of different lengths (view-group sharding: local Q, gathered K/V)
void attention_tc_kv(cudaStream_t s, const bf16* q, int ld_q, int q_col0, const bf16* kv, int ld_kv, int k_col0, int v_col0,
                     bf16* out, int batches, int seq_q, int seq_kv, int heads, int d, int dpad, float* stats = nullptr);
// `stats` (optional, fp32 [batches*seq_q, heads, 2]): the launch also reports each row's softmax reference and row sum so that
// launches over disjoint key ranges can be merged into the softmax over their union (parts in argument order)
void attention_merge(cudaStream_t s, int nparts, const bf16* const* parts, const float* const* stats, int64_t rows, int heads,
                     int dpad, bf16* out);
// norm.cu
void im2col_input(cudaStream_t s, const float* latents, int n_img, int cin, int h, int w, int kpad, bf16* out,
                  const float* premix = nullptr);
void softmax_rows(cudaStream_t s, const float* in, int64_t rows, int cols, float scale, bf16* out);
void timestep_sinusoid_bf16(cudaStream_t s, const int64_t* t, int n, int dim, bf16* out);
void groupnorm(cudaStream_t s, const bf16* x0, int c0, const bf16* x1, int c1, int n_img, int hw, int groups,
               float eps, const float* gamma, const float* beta, bool silu, bf16* out, float* scratch);
size_t groupnorm_scratch_floats(int n_img, int groups);
void groupnorm_init();  // one-time kernel attributes (call outside stream capture)
void layernorm(cudaStream_t s, const bf16* x, int rows, int c, float eps, const float* gamma, const float* beta,
               bf16* out);
void upsample_nearest2x(cudaStream_t s, const bf16* x, int n_img, int h, int w, int c, bf16* out);

// elementwise.cu
void nhwc_to_nchw_f32(cudaStream_t s, const bf16* x, int n_img, int hw, int c, float* out);
void build_inputs(cudaStream_t s, const float* x_t, const float* ctx, const float* rays, int B, int v_c, int v_t,
                  int ray_views, int ray_off, int R, int hw, float* out);
void ddpm_step(cudaStream_t s, const float* eps_c, const float* eps_u, float scale, int B, int v_c, int v_t, int chw,
               const float* x_t, const float* noise, float sa, float s1a, float c_x0, float c_xt, float sigma, float clip,
               float* x_prev);
void ddim_step(cudaStream_t s, const float* eps_c, const float* eps_u, float scale, int B, int v_c, int v_t, int chw,
               const float* x_t, float sa, float s1a, float sp, float s1p, float* x_prev, float* eps_out);
void raymap(cudaStream_t s, const float* extr, const float* intr, int n, int h, int w, bool plucker, float* out,
            int oct_o = 0, int oct_d = 0, bool srt = false);
// weight packing helpers (device side): dst bf16 [rows, ld]; all sources fp32
void convert_f32(cudaStream_t s, const void* src, int dtype, int64_t n, float* dst);
// dst[rowmap ? rowmap[r] : r, dst_col0 + j] = src[r, j]; rowmap is a device array of `rows` ints or NULL
// pack-time fp32 helpers: C = A . B and y = A . x + add
void matmul_f32(cudaStream_t s, const float* A, const float* B, float* C, int m, int n, int k);
void matvec_bias(cudaStream_t s, const float* A, const float* x, const float* add, float* y, int m, int k);
void pack_rows(cudaStream_t s, const float* src, int rows, int cols, int src_ld, bf16* dst, int dst_ld, int dst_col0,
               const int* rowmap);
void pack_conv3x3(cudaStream_t s, const float* w, int cout, int cin, int ks, bf16* dst, int dst_ld, int dst_col0);
void fill_zero(cudaStream_t s, void* p, size_t bytes);

}  // namespace mvldm
