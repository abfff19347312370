// The fused sequence kernel: ONE persistent launch (one CTA per SM, all co-resident) executes a list of ops
// back to back with a grid-wide barrier between consecutive ops instead of a kernel boundary:
//
//   SEQ_GEMM     tcgen05 / TMEM / TMA implicit GEMM: conv3x3 (stride 1/2), 1x1 conv, Linear over bf16 NHWC activations,
//                fp32 accumulation in tensor memory, fused epilogues (bias, per-image time-embedding row, residual,
//                GEGLU, SiLU, fp32-NCHW head), split-K with an in-op fixed-order reduction
//   SEQ_GN       GroupNorm(+SiLU), optionally over the channel concat of two tensors
//   SEQ_LN       LayerNorm
//   SEQ_UPSAMPLE / SEQ_IM2COL / SEQ_SINUSOID / SEQ_SPLITK_REDUCE   the small helpers of the UNet forward
//
// Why: at 1 scene x 8 views the denoiser forward is ~230 dependent kernels of 5-30 us; every boundary cost ~2 us of
// launch gap plus ~3 us of per-kernel ramp (barrier init, TMEM allocation, cold descriptors, HBM latency of the first
// weight tile).  Here the barriers, the 512 TMEM columns and the shared-memory ring are set up once per launch, an op
// boundary is one grid barrier (~1 us), and while an op drains (epilogue, reduction, barrier) the idle TMA-producer
// lane pulls the NEXT op's weight tiles into L2 (cp.async.bulk.prefetch.tensor), so the weight stream - 1.5 GB per
// forward, the whole cost of the 4x4 / 8x8 levels - no longer stops at every layer.
//
// GEMM mechanics (per CTA): D[128 x BN] (TMEM, fp32) += A[128 x 64] (smem, K-major SW128) * W[BN x 64]^T (smem).
// A is never materialised: for every (segment, filter tap, 64-channel block) one TMA box [64 ch, bw, bh, bn]
// (bw*bh*bn = 128 output pixels) is fetched from the NHWC source at the tap's pixel offset; out-of-image coordinates
// are zero-filled by the TMA unit, which is exactly the conv's zero padding.  Stride-2 convs use the tensor map's
// traversal stride.  Up to three K-segments let one accumulator take conv2(3x3) + the 1x1 shortcut over the (possibly
// concatenated) block input.  Warp roles: warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one elected lane),
// warps 2..5 = epilogue (each owns the 32 TMEM lanes of its warp%4 quarter); the accumulator is double-buffered in TMEM
// so the epilogue of work item i overlaps the main loop of item i+1.  All reductions run in a fixed order: bit-stable.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "seq.cuh"
#include "tc_common.cuh"

namespace mvldm {

CUtensorMap make_tmap_bf16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                           const uint32_t* box, const uint32_t* elem_strides) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
  });
  MV_CHECK(encode != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  CUtensorMap m;
  CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims,
                      strides_bytes, box, elem_strides, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  MV_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return m;
}

namespace {

constexpr int BM = 128, BK = 64, KC = SEQ_KC;
constexpr int A_BYTES = BM * BK * 2;  // 16 KB
constexpr int NWARPS = SEQ_THREADS / 32;
constexpr int MAX_STAGES = 8;
constexpr int CV_IMGS = 8;  // images one 128-pixel tile can span in the bias/row-vector table (hw >= 16)
constexpr int ACC_COLS = 256;
constexpr int GN_WARP_MAXV = 10;  // GroupNorm warp mode: 16-byte vectors one lane holds (x 32 lanes = vectors per item)

// ---- shared memory map (dynamic, carved by hand; base rounded up to 1024 B for the SWIZZLE_128B atoms) ----
//   [0, AUX)                     op descriptor, mbarriers, TMEM slot, reduction scratch
//   [AUX, AUX + WORK)            GEMM: `stages` ring slots of KC*(A tile + B tile), then the bias/row-vector table;
//                                GroupNorm: the CTA's channel slabs
constexpr int OFF_DESC = 0;
constexpr int OFF_BAR_FULL = 512, OFF_BAR_EMPTY = 576, OFF_ACC_FULL = 640, OFF_ACC_EMPTY = 656, OFF_TMEM = 672;
constexpr int OFF_RED = 704;    // 64 floats
constexpr int OFF_STAT = 960;   // 16 floats
constexpr int AUX_BYTES = 1024;
constexpr int WORK_BYTES = 3 * KC * (A_BYTES + 160 * 128) + CV_IMGS * 160 * 4;  // 3 stages of the 160-wide tile + its table
constexpr int SMEM_BYTES = 1024 + AUX_BYTES + WORK_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may use");

inline int stage_bytes_for(int bn) { return KC * (A_BYTES + bn * 128); }
inline int stages_for(int bn) { return std::min(MAX_STAGES, (WORK_BYTES - CV_IMGS * bn * 4) / stage_bytes_for(bn)); }

// ---- small device helpers ---------------------------------------------------------------------------------
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// erf-form GELU (F.gelu default, mvdream/attention.py:60-70) with erf from Abramowitz & Stegun 7.1.26 (|abs err| <
// 1.5e-7, far below the bf16 the result is stored in): Phi(-|x|) = 0.5 * poly(t) * exp(-x^2/2), t = 1/(1 + p |x|/sqrt2).
// 18 instructions (2 MUFU) instead of erff's ~35 with branches; the GEGLU epilogue is issue-bound on this.
__device__ __forceinline__ float gelu_exact(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float p = fmaf(t, 0.5f * 1.061405429f, 0.5f * -1.453152027f);
  p = fmaf(t, p, 0.5f * 1.421413741f);
  p = fmaf(t, p, 0.5f * -0.284496736f);
  p = fmaf(t, p, 0.5f * 0.254829592f);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float q = p * t * e;  // Phi(-|x|)
  return x * (x < 0.f ? q : 1.f - q);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t w) {
  return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w));
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y), c = unpack_bf16(u.z), d = unpack_bf16(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
  return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
// activations are written by other CTAs of the SAME launch: read them through L2 only (no stale L1 line can be hit)
__device__ __forceinline__ uint4 ld_cg16(const void* p) { return __ldcg(reinterpret_cast<const uint4*>(p)); }

// 256-bit global store: one instruction covers a full 32-byte sector per thread (rows are >= 64 B apart, so 16-byte
// stores would touch every sector twice)
__device__ __forceinline__ void st_global_v8(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                             uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ long long globaltimer_ns() {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// timeline of CTA 0 for the per-op report: (globaltimer ns, clock64) pairs
__device__ __forceinline__ void stamp(long long* timing, unsigned k) {
  timing[2 * k] = globaltimer_ns();
  timing[2 * k + 1] = clock64();
}

// ---- grid-wide barrier -----------------------------------------------------------------------------------
// All CTAs of the launch are co-resident (cooperative launch, one CTA per SM).  A monotonic arrival counter: barrier k
// completes when it reaches k * gridDim.x.  Every thread publishes its global writes (gpu-scope fence; the proxy fence
// orders them against the TMA = async-proxy reads other CTAs will issue), thread 0 arrives with release semantics and
// polls with acquire semantics, the CTA barrier on either side extends that to the whole CTA (the cooperative-groups
// grid.sync construction).  While thread 0 polls, warp 1 stages the next op's descriptor in shared memory.
struct GridBarrier {
  unsigned* ctr;
  unsigned n;  // barriers passed
  long long* timing;
  __device__ __forceinline__ void sync(const SeqOp* next, uint8_t* desc_smem) {
    __threadfence();
    asm volatile("fence.proxy.async.global;" ::: "memory");
    __syncthreads();
    ++n;
    if (threadIdx.x == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
    if (next != nullptr && threadIdx.x >= 32 && threadIdx.x < 32 + SEQ_OPC_BYTES / 16)
      reinterpret_cast<uint4*>(desc_smem)[threadIdx.x - 32] = __ldg(reinterpret_cast<const uint4*>(next) + (threadIdx.x - 32));
    if (threadIdx.x == 0) {
      const unsigned target = n * gridDim.x;
      unsigned spins = 0;
      while (ld_acquire_u32(ctr) < target) {
        if (++spins > (1u << 25)) __trap();  // a protocol bug / lost co-residency traps instead of hanging the GPU
      }
      if (timing != nullptr && blockIdx.x == 0) stamp(timing, n);
    }
    __syncthreads();
    asm volatile("fence.proxy.async.global;" ::: "memory");
  }
};

// role state that survives from op to op: mbarrier phase parities per ring slot, accumulator-buffer counter
struct RoleState {
  uint32_t empty_par;  // producer: parity to wait for on bar_empty[s] before refilling slot s
  uint32_t full_par;   // MMA issuer: parity to wait for on bar_full[s]
  uint32_t acc_n;      // MMA issuer / epilogue: accumulator hand-offs so far (buffer = acc_n & 1, phase = (acc_n >> 1) & 1)
};

// The tensor maps live in global memory (written by cudaMemcpy before the launch) and are read through the tensormap proxy:
// each CTA acquires them once before first use (CUDA programming guide, "tensor map in global memory").  Done one op ahead.
__device__ __forceinline__ void acquire_tensormaps(const SeqOp* o, int nseg) {
  for (int s = 0; s < nseg; ++s)
    for (int k = 0; k < KC; ++k)
      asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(&o->tmA[s][k])) : "memory");
  if (nseg > 0)
    for (int k = 0; k < KC; ++k)
      asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(&o->tmB[k])) : "memory");
}

// L2 prefetch of the next GEMM op's weight tiles this CTA will consume
__device__ __forceinline__ void prefetch_next_weights(const SeqPrefetch& pf) {
  if (pf.map == nullptr) return;
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(pf.map)) : "memory");
  const int work = pf.mt * pf.nt * pf.splits;
  for (int w = blockIdx.x; w < work; w += gridDim.x) {
    if (w % pf.mt != 0) continue;  // the m-tiles of one (n-tile, split) share the weight tile: fetched once
    const int ntile = (w / pf.mt) % pf.nt, z = w / (pf.mt * pf.nt);
    const int c_end = pf.chunk0[z + 1];
    for (int ch = pf.chunk0[z]; ch < c_end; ch += KC)
      asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(
                       reinterpret_cast<uint64_t>(pf.map)),
                   "r"(0), "r"(ntile * pf.bn), "r"(ch)
                   : "memory");
  }
}

// =====================================================================================================
// SEQ_GEMM
// =====================================================================================================
__device__ __noinline__ void gemm_op(const SeqGemm& g, const SeqPrefetch& pf, int next_nseg, const SeqOp* opg, uint8_t* smem,
                                        uint32_t smem_base, uint32_t tmem_base, RoleState& st) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.bn;
  const uint32_t b_bytes = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = KC * (A_BYTES + b_bytes);
  const uint32_t ring_base = smem_base + AUX_BYTES;
  const uint32_t bar_full = smem_base + OFF_BAR_FULL, bar_empty = smem_base + OFF_BAR_EMPTY;
  const uint32_t bar_acc_full = smem_base + OFF_ACC_FULL, bar_acc_empty = smem_base + OFF_ACC_EMPTY;
  const int num_work = g.mt * g.nt * g.splits;

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int ring = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int mtile = w % g.mt, ntile = (w / g.mt) % g.nt, z = w / (g.mt * g.nt);
        const int m0 = mtile * BM, n0 = ntile * BN;
        const int st_begin = z * g.steps_per_split;
        const int nst = min(g.num_steps - st_begin, g.steps_per_split);
        const int img0 = m0 / g.hw;
        const int y0 = (m0 - img0 * g.hw) / g.ow;
        // locate (segment, tap, channel block) of the first step
        int s = 0, t = 0, cb = st_begin;
        while (cb >= g.seg[s].ntaps * g.seg[s].spt) {
          cb -= g.seg[s].ntaps * g.seg[s].spt;
          ++s;
        }
        t = cb / g.seg[s].spt;
        cb = (cb - t * g.seg[s].spt) * KC;
        for (int i = 0; i < nst; ++i) {
          const SeqSeg& sg = g.seg[s];
          const int kc = min(KC, sg.ncblk - cb);
          tc::mbar_wait(bar_empty + 8 * ring, (st.empty_par >> ring) & 1u);
          st.empty_par ^= 1u << ring;
          const uint32_t full = bar_full + 8 * ring;
          tc::mbar_expect_tx(full, kc * (A_BYTES + b_bytes));
          const uint32_t sa = ring_base + ring * stage_bytes;
          tc::tma_load_5d(sa, &opg->tmA[s][kc - 1], full, 0, sg.dw[t], y0 * sg.stride + sg.dh[t], img0, sg.cblk[t] + cb);
          tc::tma_load_3d(sa + KC * A_BYTES, &opg->tmB[kc - 1], full, 0, n0, sg.kchunk0 + t * sg.ncblk + cb);
          ring = ring + 1 == g.stages ? 0 : ring + 1;
          cb += kc;
          if (cb == sg.ncblk) {
            cb = 0;
            if (++t == sg.ntaps) {
              t = 0;
              ++s;
            }
          }
        }
      }
      acquire_tensormaps(opg + 1, next_nseg);
      prefetch_next_weights(pf);  // the rest of this op (MMA tail, epilogue, reduction, barrier) hides the HBM latency
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // the whole warp runs the warp-uniform loop and the barrier waits; one elected lane issues tcgen05.mma / commit,
    // which lets ptxas emit the UTCHMMAs back to back instead of one ELECT/branch loop per instruction
    const uint32_t idesc = tc::umma_idesc_bf16(BM, BN, false, false);
    int ring = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int z = w / (g.mt * g.nt);
      const int st_begin = z * g.steps_per_split;
      const int nst = min(g.num_steps - st_begin, g.steps_per_split);
      int s = 0, t = 0, cb = st_begin;  // same walk as the producer, to know how many chunks each step carries
      while (cb >= g.seg[s].ntaps * g.seg[s].spt) {
        cb -= g.seg[s].ntaps * g.seg[s].spt;
        ++s;
      }
      t = cb / g.seg[s].spt;
      cb = (cb - t * g.seg[s].spt) * KC;
      const uint32_t ab = st.acc_n & 1u;
      tc::mbar_wait(bar_acc_empty + 8 * ab, ((st.acc_n >> 1) & 1u) ^ 1u);  // epilogue has drained this buffer
      tc::tc_fence_after();
      const uint32_t tmem_d = tmem_base + ab * ACC_COLS;
      for (int i = 0; i < nst; ++i) {
        const int kc = min(KC, (int)g.seg[s].ncblk - cb);
        tc::mbar_wait(bar_full + 8 * ring, (st.full_par >> ring) & 1u);
        st.full_par ^= 1u << ring;
        tc::tc_fence_after();
        const uint32_t sa = ring_base + ring * stage_bytes;
        if (tc::elect_one()) {
          for (int c = 0; c < kc; ++c) {
            const uint64_t adesc = tc::umma_desc_k_sw128(sa + c * A_BYTES);
            const uint64_t bdesc = tc::umma_desc_k_sw128(sa + KC * A_BYTES + c * b_bytes);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k)  // +32 bytes (=2 in descriptor units) per K=16 slice inside the swizzle atom
              tc::umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (i | c | k) != 0);
          }
          tc::umma_commit(bar_empty + 8 * ring);  // frees the smem slot when these MMAs retire
        }
        __syncwarp();
        ring = ring + 1 == g.stages ? 0 : ring + 1;
        cb += kc;
        if (cb == g.seg[s].ncblk) {
          cb = 0;
          if (++t == g.seg[s].ntaps) {
            t = 0;
            ++s;
          }
        }
      }
      if (tc::elect_one()) tc::umma_commit(bar_acc_full + 8 * ab);
      __syncwarp();
      ++st.acc_n;
    }
  } else if (warp < 6) {
    // ================= epilogue =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..127 among the epilogue threads
    float* colvec = reinterpret_cast<float*>(smem + AUX_BYTES + g.stages * stage_bytes);  // [image in tile][BN]
    for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
      const int mtile = w % g.mt, ntile = (w / g.mt) % g.nt, z = w / (g.mt * g.nt);
      const int m0 = mtile * BM, n0 = ntile * BN;
      const uint32_t ab = st.acc_n & 1u;
      const uint32_t tmem_d = tmem_base + ab * ACC_COLS;
      const int m = m0 + row;
      const bool ok = m < g.M;
      const int img = m / g.hw;
      // ---- while the main loop of this item runs: stage everything the epilogue needs that is not the accumulator
      const int img0 = m0 / g.hw;
      const int imgs_in_tile = (BM + g.hw - 1) / g.hw;
      const bool use_table = (!g.partial || g.counters) && (g.bias || g.rowvec) && imgs_in_tile <= CV_IMGS;
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous item's readers of the table are done
      if (use_table) {
        for (int i = et; i < imgs_in_tile * BN; i += 128) {
          const int b = i / BN, c = i - b * BN;
          float v = g.bias ? g.bias[n0 + c] : 0.f;
          if (g.rowvec && (int64_t)(img0 + b) * g.hw < g.M) v += __ldcg(g.rowvec + (int64_t)(img0 + b) * g.rowvec_ld + n0 + c);
          colvec[b * BN + c] = v;
        }
      }
      // residual (bf16 row of this thread): chunk c+1 is fetched while chunk c is converted and stored
      const bool use_res = ok && !g.partial && g.mode == 0 && g.residual;
      const uint4* res_row = use_res ? reinterpret_cast<const uint4*>(g.residual + (int64_t)m * g.res_ld + n0) : nullptr;
      uint4 res_next[4];
      if (use_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) res_next[j] = __ldcg(res_row + j);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float* cvrow = colvec + (img - img0) * BN;
      tc::mbar_wait(bar_acc_full + 8 * ab, (st.acc_n >> 1) & 1u);
      tc::tc_fence_after();
#pragma unroll 1  // rolled: the unrolled epilogue (x8 chunks x 3 modes) cost 0.5 ms per forward in code size / registers
      for (int c0 = 0; c0 < BN; c0 += 32) {
        uint32_t r[32];
        uint4 res_cur[4];
        if (use_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) res_cur[j] = res_next[j];
          if (c0 + 32 < BN) {
#pragma unroll
            for (int j = 0; j < 4; ++j) res_next[j] = __ldcg(res_row + (c0 + 32) / 8 + j);
          }
        }
        __syncwarp();
        tc::tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
        tc::tmem_ld_wait();
        if (ok && g.partial) {  // split-K: raw fp32 partial, reduced (+ epilogue) below or by the SEQ_SPLITK_REDUCE op
          float* pp = g.partial + ((int64_t)z * g.M + m) * g.N + n0 + c0;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_global_v8(pp + 8 * j, r[8 * j], r[8 * j + 1], r[8 * j + 2], r[8 * j + 3], r[8 * j + 4], r[8 * j + 5], r[8 * j + 6],
                         r[8 * j + 7]);
        } else if (ok) {
          const int n = n0 + c0;
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          if (use_table) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 b = *reinterpret_cast<const float4*>(cvrow + c0 + j);  // smem, same address across the warp's rows of one image
              v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
            }
          } else {
            if (g.bias) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = *reinterpret_cast<const float4*>(g.bias + n + j);
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
            if (g.rowvec) {
              const float* rv = g.rowvec + (int64_t)img * g.rowvec_ld + n;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float4 b = __ldcg(reinterpret_cast<const float4*>(rv + j));
                v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
              }
            }
          }
          if (g.mode == 0 || g.mode == 3) {
            if (use_res) {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint4 u = res_cur[j];
                const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float2 f = unpack_bf16(wd[e]);
                  v[j * 8 + e * 2] += f.x;
                  v[j * 8 + e * 2 + 1] += f.y;
                }
              }
            }
            if (g.mode == 3) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
            }
            bf16* op = reinterpret_cast<bf16*>(g.out) + (int64_t)m * g.ldo + n;
#pragma unroll
            for (int j = 0; j < 2; ++j)
              st_global_v8(op + 16 * j, pack_bf16(v[j * 16], v[j * 16 + 1]), pack_bf16(v[j * 16 + 2], v[j * 16 + 3]),
                           pack_bf16(v[j * 16 + 4], v[j * 16 + 5]), pack_bf16(v[j * 16 + 6], v[j * 16 + 7]),
                           pack_bf16(v[j * 16 + 8], v[j * 16 + 9]), pack_bf16(v[j * 16 + 10], v[j * 16 + 11]),
                           pack_bf16(v[j * 16 + 12], v[j * 16 + 13]), pack_bf16(v[j * 16 + 14], v[j * 16 + 15]));
          } else if (g.mode == 1) {
            // columns [0,16) = values, [16,32) = gates of the same 16 hidden channels
            float gl[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) gl[j] = v[j] * gelu_exact(v[16 + j]);
            st_global_v8(reinterpret_cast<bf16*>(g.out) + (int64_t)m * g.ldo + n / 2, pack_bf16(gl[0], gl[1]),
                         pack_bf16(gl[2], gl[3]), pack_bf16(gl[4], gl[5]), pack_bf16(gl[6], gl[7]), pack_bf16(gl[8], gl[9]),
                         pack_bf16(gl[10], gl[11]), pack_bf16(gl[12], gl[13]), pack_bf16(gl[14], gl[15]));
          } else if (g.mode == 4) {
            float* op = reinterpret_cast<float*>(g.out) + (int64_t)m * g.ldo + n;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              st_global_v8(op + 8 * j, __float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]), __float_as_uint(v[8 * j + 2]),
                           __float_as_uint(v[8 * j + 3]), __float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]),
                           __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]));
          } else {
            const int pix = m - img * g.hw;
            float* op = reinterpret_cast<float*>(g.out) + (int64_t)img * g.n_valid * g.hw + pix;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n + j < g.n_valid) op[(int64_t)(n + j) * g.hw] = v[j];
          }
        }
      }
      tc::tc_fence_before();
      tc::mbar_arrive(bar_acc_empty + 8 * ab);  // this thread is done reading the accumulator buffer
      ++st.acc_n;
      if (g.counters) {
        // ---- split-K reduction fused into the op: all splits of a tile are co-resident (one work item per CTA), so
        // they meet at a global counter; each then reduces 1/splits of the tile's rows in fixed z order (bit-stable)
        // and applies the epilogue.
        const int tile = mtile + ntile * g.mt;
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {
          atomicAdd(&g.counters[tile], 1);
          uint32_t spins = 0;
          while (*reinterpret_cast<volatile int*>(&g.counters[tile]) < g.splits) {
            if (++spins > (1u << 25)) __trap();
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        __threadfence();
        const int rows_per = (BM + g.splits - 1) / g.splits;
        const int r0 = z * rows_per, r1 = min(BM, r0 + rows_per);
        const int NV = BN / 8;
        const int items = (r1 - r0) * NV;
        const int64_t zstride = (int64_t)g.M * g.N;
        for (int i0 = et; i0 < items; i0 += 128 * 4) {
          // 4 independent items per thread.  Dead slots (past the end / past M) alias the thread's first item so that
          // every load below is unconditional and the 4 x 2 x splits requests are all in flight together; only the
          // final store is predicated.
          int mmv[4], nnv[4];
          bool live[4];
          float v[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 128;
            const int mm = m0 + r0 + i / NV;
            live[u] = i < items && mm < g.M;
            const int ii = live[u] ? i : i0;
            mmv[u] = min(m0 + r0 + ii / NV, g.M - 1);
            nnv[u] = (ii % NV) * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
          }
          uint4 rres[4];
          float4 cv0[4], cv1[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            rres[u] = g.residual ? ld_cg16(g.residual + (int64_t)mmv[u] * g.res_ld + n0 + nnv[u]) : make_uint4(0u, 0u, 0u, 0u);
            if (use_table) {
              const float* cr = colvec + (mmv[u] / g.hw - img0) * BN + nnv[u];
              cv0[u] = *reinterpret_cast<const float4*>(cr);
              cv1[u] = *reinterpret_cast<const float4*>(cr + 4);
            } else {
              float t8[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                t8[j] = g.bias ? g.bias[n0 + nnv[u] + j] : 0.f;
                if (g.rowvec) t8[j] += __ldcg(g.rowvec + (int64_t)(mmv[u] / g.hw) * g.rowvec_ld + n0 + nnv[u] + j);
              }
              cv0[u] = make_float4(t8[0], t8[1], t8[2], t8[3]);
              cv1[u] = make_float4(t8[4], t8[5], t8[6], t8[7]);
            }
          }
#pragma unroll 4
          for (int zz = 0; zz < g.splits; ++zz) {  // fixed z order: bit-stable
            float4 a[4], b4[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float4* pp = reinterpret_cast<const float4*>(g.partial + zz * zstride + (int64_t)mmv[u] * g.N + n0 + nnv[u]);
              a[u] = __ldcg(pp);
              b4[u] = __ldcg(pp + 1);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              v[u][0] += a[u].x; v[u][1] += a[u].y; v[u][2] += a[u].z; v[u][3] += a[u].w;
              v[u][4] += b4[u].x; v[u][5] += b4[u].y; v[u][6] += b4[u].z; v[u][7] += b4[u].w;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            v[u][0] += cv0[u].x; v[u][1] += cv0[u].y; v[u][2] += cv0[u].z; v[u][3] += cv0[u].w;
            v[u][4] += cv1[u].x; v[u][5] += cv1[u].y; v[u][6] += cv1[u].z; v[u][7] += cv1[u].w;
            const uint32_t wds[4] = {rres[u].x, rres[u].y, rres[u].z, rres[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = unpack_bf16(wds[e]);
              v[u][2 * e] += f.x;
              v[u][2 * e + 1] += f.y;
            }
            if (live[u])
              *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(g.out) + (int64_t)mmv[u] * g.ldo + n0 + nnv[u]) = pack8(v[u]);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) {  // the last split to finish re-arms the counters for the next op
          if (atomicAdd(&g.counters[g.mt * g.nt + tile], 1) == g.splits - 1) {
            g.counters[tile] = 0;
            g.counters[g.mt * g.nt + tile] = 0;
          }
        }
      }
    }
  }
  // (warps 6, 7 have no GEMM role: they wait at the op barrier)
}

// =====================================================================================================
// SEQ_GN  GroupNorm (+SiLU), two-pass (mean, then centred second moment) so that |mean| >> std costs no precision
// =====================================================================================================
// block-wide sums of up to 4 per-group values: fixed shuffle tree inside a warp, warps folded in order (bit-stable)
__device__ __forceinline__ void block_sum4(float (&v)[4], float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int g = 0; g < 4; ++g) v[g] = warp_sum(v[g]);
  __syncthreads();  // previous readers of `red` are done
  if (lane == 0) {
#pragma unroll
    for (int g = 0; g < 4; ++g) red[warp * 4 + g] = v[g];
  }
  __syncthreads();
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float a = 0.f;
#pragma unroll
    for (int w = 0; w < NWARPS; ++w) a += red[w * 4 + g];
    v[g] = a;
  }
}

// One CTA per (image, channel block, pixel chunk).  Thread t owns vector column cv = t % NV (8 channels, at most two
// groups) and walks pixels pl, pl + lanes_p, ...; the chunk is cached in shared memory so the centred second pass and the
// normalisation never go back to L2.  With ps > 1 chunk statistics (mean, M2) meet in global memory across one grid
// barrier and are merged with Chan's formula in chunk order.
__device__ __noinline__ void gn_cta(const SeqGN& p, uint8_t* smem, GridBarrier& gb) {
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  float* stat = reinterpret_cast<float*>(smem + OFF_STAT);  // [4][2]: mean, rstd
  bf16* slab0 = reinterpret_cast<bf16*>(smem + AUX_BYTES);
  const int C = p.c0 + p.c1, NV = p.cb / 8, GB = p.cb / p.cgn, nblk = C / p.cb;
  const int t = threadIdx.x, lanes_p = SEQ_THREADS / NV, cv = t % NV, pl = t / NV;
  const bool active = pl < lanes_p;
  const int g_lo = (cv * 8) / p.cgn, eb = (g_lo + 1) * p.cgn - cv * 8;  // channels e < eb of the vector -> g_lo, the rest g_lo + 1
  const int slab_elems = p.px * p.cb;
  const float inv_cnt = 1.f / ((float)p.px * (float)p.cgn);

  auto locate = [&](int item, const bf16*& src, int64_t& pitch, bf16*& dst) {
    const int chunk = item % p.ps, blk = (item / p.ps) % nblk, img = item / (p.ps * nblk);
    const int ch = blk * p.cb + cv * 8;
    const int64_t pix0 = (int64_t)img * p.hw + (int64_t)chunk * p.px;
    const bool first = ch < p.c0;
    src = first ? p.x0 + pix0 * p.c0 + ch : p.x1 + pix0 * p.c1 + (ch - p.c0);
    pitch = first ? p.c0 : p.c1;
    dst = p.out + pix0 * C + ch;
  };
  auto apply = [&](const bf16* src, int64_t pitch, bf16* dst, const bf16* slab, int blk, float m_lo, float r_lo, float m_hi,
                   float r_hi) {
    const int ch = blk * p.cb + cv * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + ch), g1 = *reinterpret_cast<const float4*>(p.gamma + ch + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(p.beta + ch), b1 = *reinterpret_cast<const float4*>(p.beta + ch + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const bool lo = e < eb;
      sc[e] = (lo ? r_lo : r_hi) * gg[e];
      sh[e] = bb[e] - (lo ? m_lo : m_hi) * sc[e];
    }
#pragma unroll 4
    for (int pp = pl; pp < p.px; pp += lanes_p) {
      const uint4 u = p.cache ? *reinterpret_cast<const uint4*>(slab + (int64_t)pp * p.cb + cv * 8) : ld_cg16(src + pp * pitch);
      float f[8];
      unpack8(u, f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = fmaf(f[e], sc[e], sh[e]);
        f[e] = p.silu ? silu_f(y) : y;
      }
      *reinterpret_cast<uint4*>(dst + (int64_t)pp * C) = pack8(f);
    }
  };

  int k = 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++k) {
    const bf16* src;
    int64_t pitch;
    bf16* dst;
    locate(item, src, pitch, dst);
    bf16* slab = slab0 + (int64_t)k * slab_elems;
    // ---- pass 1: load (and cache) the chunk, per-channel sums -> group means
    float sa[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) sa[e] = 0.f;
    if (active) {
      int pp = pl;
      for (; pp + 7 * lanes_p < p.px; pp += 8 * lanes_p) {  // 8 independent 16-byte loads in flight per thread
        uint4 u[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) u[j] = ld_cg16(src + (int64_t)(pp + j * lanes_p) * pitch);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (p.cache) *reinterpret_cast<uint4*>(slab + (int64_t)(pp + j * lanes_p) * p.cb + cv * 8) = u[j];
          float f[8];
          unpack8(u[j], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) sa[e] += f[e];
        }
      }
      for (; pp < p.px; pp += lanes_p) {
        const uint4 u = ld_cg16(src + (int64_t)pp * pitch);
        if (p.cache) *reinterpret_cast<uint4*>(slab + (int64_t)pp * p.cb + cv * 8) = u;
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) sa[e] += f[e];
      }
    }
    float S[4];
    {
      float s_lo = 0.f, s_hi = 0.f;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s_lo += e < eb ? sa[e] : 0.f;
        s_hi += e < eb ? 0.f : sa[e];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) S[g] = active ? ((g == g_lo ? s_lo : 0.f) + (g == g_lo + 1 ? s_hi : 0.f)) : 0.f;
    }
    block_sum4(S, red);
    const int g_hi = min(g_lo + 1, GB - 1);
    float mean_lo = 0.f, mean_hi = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      mean_lo = g == g_lo ? S[g] * inv_cnt : mean_lo;
      mean_hi = g == g_hi ? S[g] * inv_cnt : mean_hi;
    }
    // ---- pass 2: centred second moment
    float q_lo = 0.f, q_hi = 0.f;
    if (active) {
#pragma unroll 4
      for (int pp = pl; pp < p.px; pp += lanes_p) {
        const uint4 u = p.cache ? *reinterpret_cast<const uint4*>(slab + (int64_t)pp * p.cb + cv * 8) : ld_cg16(src + pp * pitch);
        float f[8];
        unpack8(u, f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool lo = e < eb;
          const float d = f[e] - (lo ? mean_lo : mean_hi);
          q_lo = lo ? fmaf(d, d, q_lo) : q_lo;
          q_hi = lo ? q_hi : fmaf(d, d, q_hi);
        }
      }
    }
    float Q[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) Q[g] = (g == g_lo ? q_lo : 0.f) + (g == g_lo + 1 ? q_hi : 0.f);
    block_sum4(Q, red);
    if (p.ps == 1) {
      float r_lo = 0.f, r_hi = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float r = rsqrtf(Q[g] * inv_cnt + p.eps);
        r_lo = g == g_lo ? r : r_lo;
        r_hi = g == g_hi ? r : r_hi;
      }
      if (active) apply(src, pitch, dst, slab, (item / p.ps) % nblk, mean_lo, r_lo, mean_hi, r_hi);
    } else if (t < GB) {
      float mg = 0.f, qg = 0.f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        mg = g == t ? S[g] * inv_cnt : mg;
        qg = g == t ? Q[g] : qg;
      }
      *reinterpret_cast<float2*>(p.partial + ((int64_t)item * 4 + t) * 2) = make_float2(mg, qg);
    }
  }
  if (p.ps == 1) return;
  gb.sync(nullptr, nullptr);
  k = 0;
  for (int item = blockIdx.x; item < p.n_items; item += gridDim.x, ++k) {
    const bf16* src;
    int64_t pitch;
    bf16* dst;
    locate(item, src, pitch, dst);
    __syncthreads();  // previous item's readers of `stat` are done
    if (t < GB) {     // merge the image's chunks in chunk order (equal counts): mean of means, M2 += n (mean_c - mean)^2
      const int64_t first = (int64_t)(item - item % p.ps);
      float msum = 0.f;
      for (int c = 0; c < p.ps; ++c) msum += __ldcg(p.partial + ((first + c) * 4 + t) * 2);
      const float mean = msum / (float)p.ps;
      float m2 = 0.f;
      const float n_c = (float)p.px * (float)p.cgn;
      for (int c = 0; c < p.ps; ++c) {
        const float2 pr = __ldcg(reinterpret_cast<const float2*>(p.partial + ((first + c) * 4 + t) * 2));
        const float d = pr.x - mean;
        m2 += pr.y + n_c * d * d;
      }
      stat[2 * t] = mean;
      stat[2 * t + 1] = rsqrtf(m2 / (n_c * (float)p.ps) + p.eps);
    }
    __syncthreads();
    const int g_hi = min(g_lo + 1, GB - 1);
    if (active)
      apply(src, pitch, dst, slab0 + (int64_t)k * slab_elems, (item / p.ps) % nblk, stat[2 * g_lo], stat[2 * g_lo + 1], stat[2 * g_hi],
            stat[2 * g_hi + 1]);
  }
}

// Small images (<= 320 vectors per (image, channel block)): one WARP per item, the block in registers, shuffle-only
// reductions, no CTA barrier; the 8 warps of a CTA work on 8 items at once.
__device__ __noinline__ void gn_warp(const SeqGN& p) {
  constexpr int MAXV = GN_WARP_MAXV;
  const int C = p.c0 + p.c1, NV = p.cb / 8, GB = p.cb / p.cgn, nblk = C / p.cb;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * NWARPS + warp, GW = gridDim.x * NWARPS;
  const int nvec = p.hw * NV;
  const float inv_cnt = 1.f / ((float)p.hw * (float)p.cgn);
  for (int item = gw; item < p.n_items; item += GW) {
    const int blk = item % nblk, img = item / nblk;
    uint4 raw[MAXV];
    int cvs[MAXV];  // vector column of slot j (NV is not a power of two: computed once)
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int v = lane + 32 * j;
      const int pix = (int)__umulhi((unsigned)v, p.nv_magic);
      cvs[j] = v - pix * NV;
      raw[j] = make_uint4(0u, 0u, 0u, 0u);
      if (v < nvec) {
        const int ch = blk * p.cb + cvs[j] * 8;
        const int64_t pixel = (int64_t)img * p.hw + pix;
        raw[j] = ld_cg16(ch < p.c0 ? p.x0 + pixel * p.c0 + ch : p.x1 + pixel * p.c1 + (ch - p.c0));
      }
    }
    float S[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int g_lo = (cvs[j] * 8) / p.cgn, eb = (g_lo + 1) * p.cgn - cvs[j] * 8;
      float f[8], s_lo = 0.f, s_hi = 0.f;
      unpack8(raw[j], f);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s_lo += e < eb ? f[e] : 0.f;
        s_hi += e < eb ? 0.f : f[e];
      }
#pragma unroll
      for (int g = 0; g < 4; ++g) S[g] += (g == g_lo ? s_lo : 0.f) + (g == g_lo + 1 ? s_hi : 0.f);  // empty slots add zeros
    }
    float mean[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) mean[g] = warp_sum(S[g]) * inv_cnt;
    float Q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      if (lane + 32 * j < nvec) {
        const int g_lo = (cvs[j] * 8) / p.cgn, eb = (g_lo + 1) * p.cgn - cvs[j] * 8;
        float m_lo = 0.f, m_hi = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          m_lo = g == g_lo ? mean[g] : m_lo;
          m_hi = g == g_lo + 1 ? mean[g] : m_hi;
        }
        float f[8], q_lo = 0.f, q_hi = 0.f;
        unpack8(raw[j], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool lo = e < eb;
          const float d = f[e] - (lo ? m_lo : m_hi);
          q_lo = lo ? fmaf(d, d, q_lo) : q_lo;
          q_hi = lo ? q_hi : fmaf(d, d, q_hi);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) Q[g] += (g == g_lo ? q_lo : 0.f) + (g == g_lo + 1 ? q_hi : 0.f);
      }
    }
    float rstd[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) rstd[g] = rsqrtf(warp_sum(Q[g]) * inv_cnt + p.eps);
    (void)GB;
#pragma unroll
    for (int j = 0; j < MAXV; ++j) {
      const int v = lane + 32 * j;
      if (v < nvec) {
        const int pix = (int)__umulhi((unsigned)v, p.nv_magic);
        const int g_lo = (cvs[j] * 8) / p.cgn, eb = (g_lo + 1) * p.cgn - cvs[j] * 8;
        float m_lo = 0.f, m_hi = 0.f, r_lo = 0.f, r_hi = 0.f;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          m_lo = g == g_lo ? mean[g] : m_lo;
          r_lo = g == g_lo ? rstd[g] : r_lo;
          m_hi = g == g_lo + 1 ? mean[g] : m_hi;
          r_hi = g == g_lo + 1 ? rstd[g] : r_hi;
        }
        const int ch = blk * p.cb + cvs[j] * 8;
        const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + ch), g1 = *reinterpret_cast<const float4*>(p.gamma + ch + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(p.beta + ch), b1 = *reinterpret_cast<const float4*>(p.beta + ch + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        float f[8];
        unpack8(raw[j], f);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const bool lo = e < eb;
          const float y = (f[e] - (lo ? m_lo : m_hi)) * (lo ? r_lo : r_hi) * gg[e] + bb[e];
          f[e] = p.silu ? silu_f(y) : y;
        }
        *reinterpret_cast<uint4*>(p.out + ((int64_t)img * p.hw + pix) * C + ch) = pack8(f);
      }
    }
  }
}

// =====================================================================================================
// SEQ_LN  LayerNorm over the channel dim: one warp per token, R tokens in flight per warp, rows in registers
// =====================================================================================================
template <int MAXV>
__device__ __noinline__ void ln_rows(const SeqLN& p) {
  constexpr int R = MAXV >= 8 ? 2 : 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * NWARPS + warp, GW = gridDim.x * NWARPS;
  const int nv = p.c / 8;
  const float inv_c = 1.f / (float)p.c;
  for (int r0 = gw; r0 < p.rows; r0 += R * GW) {
    uint4 raw[R][MAXV];
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int row = r0 + u * GW;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + 32 * i;
        raw[u][i] = (row < p.rows && v < nv) ? ld_cg16(p.x + (int64_t)row * p.c + v * 8) : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < R; ++u) {
      const int row = r0 + u * GW;
      if (row >= p.rows) break;  // warp-uniform
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        float f[8];
        unpack8(raw[u][i], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[j];
      }
      const float mean = warp_sum(s) * inv_c;
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        if (lane + 32 * i < nv) {
          float f[8];
          unpack8(raw[u][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float d = f[j] - mean;
            q = fmaf(d, d, q);
          }
        }
      }
      const float rstd = rsqrtf(warp_sum(q) * inv_c + p.eps);
#pragma unroll
      for (int i = 0; i < MAXV; ++i) {
        const int v = lane + 32 * i;
        if (v < nv) {
          const float4 g0 = *reinterpret_cast<const float4*>(p.gamma + v * 8), g1 = *reinterpret_cast<const float4*>(p.gamma + v * 8 + 4);
          const float4 b0 = *reinterpret_cast<const float4*>(p.beta + v * 8), b1 = *reinterpret_cast<const float4*>(p.beta + v * 8 + 4);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float f[8];
          unpack8(raw[u][i], f);
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * gg[j] + bb[j];
          *reinterpret_cast<uint4*>(p.out + (int64_t)row * p.c + v * 8) = pack8(f);
        }
      }
    }
  }
}

// =====================================================================================================
// small elementwise ops
// =====================================================================================================
__device__ __noinline__ void upsample_op(const SeqEW& p) {  // nearest 2x, bf16 NHWC
  const bf16* x = reinterpret_cast<const bf16*>(p.src);
  bf16* out = reinterpret_cast<bf16*>(p.dst);
  const int cv = p.c / 8;
  const int64_t total = (int64_t)p.n_img * 4 * p.h * p.w * cv;
  for (int64_t i = (int64_t)blockIdx.x * SEQ_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * SEQ_THREADS) {
    const int v = (int)(i % cv);
    int64_t q = i / cv;
    const int ox = (int)(q % (2 * p.w));
    q /= 2 * p.w;
    const int oy = (int)(q % (2 * p.h));
    const int img = (int)(q / (2 * p.h));
    *reinterpret_cast<uint4*>(out + i * 8) = ld_cg16(x + (((int64_t)img * p.h + oy / 2) * p.w + ox / 2) * p.c + v * 8);
  }
}

// conv_in operand: explicit im2col of the fp32 NCHW denoiser input.
// out[m, tap*cin + c] = latents[img, c, y+r-1, x+s-1] (zero outside / for k >= 9*cin), m = (img, y, x)
__device__ __noinline__ void im2col_op(const SeqEW& p) {
  const float* x = reinterpret_cast<const float*>(p.src);
  bf16* out = reinterpret_cast<bf16*>(p.dst);
  const int cin = p.c, h = p.h, w = p.w, kpad = p.aux;
  const int kv = kpad / 8;  // one thread = eight consecutive k of one output pixel = one 16-byte store
  const int total = p.n_img * h * w * kv;
  for (int i = blockIdx.x * SEQ_THREADS + threadIdx.x; i < total; i += gridDim.x * SEQ_THREADS) {
    const int m = i / kv, k0 = (i - m * kv) * 8;
    const int px = m % w, py = (m / w) % h, img = m / (w * h);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      v[j] = 0.f;
      if (k < 9 * cin) {
        const int tap = k / cin, c = k - tap * cin;
        const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) v[j] = __ldcg(x + ((img * cin + c) * h + yy) * w + xx);
      }
    }
    *reinterpret_cast<uint4*>(out + (int64_t)i * 8) = pack8(v);
  }
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin] -> bf16 (the operand autocast feeds
// time_embedding.linear_1 in the reference)
__device__ __noinline__ void sinusoid_op(const SeqEW& p) {
  const int64_t* t = reinterpret_cast<const int64_t*>(p.src);
  bf16* out = reinterpret_cast<bf16*>(p.dst);
  const int n = p.n_img, dim = p.c, half = dim / 2;
  for (int i = blockIdx.x * SEQ_THREADS + threadIdx.x; i < n * half; i += gridDim.x * SEQ_THREADS) {
    const int r = i / half, j = i - r * half;
    const float freq = expf(-logf(10000.f) * (float)j / (float)half);
    const float arg = (float)__ldcg(t + r) * freq;
    out[(int64_t)r * dim + j] = (bf16)cosf(arg);
    out[(int64_t)r * dim + half + j] = (bf16)sinf(arg);
  }
}

// split-K second pass (tiles x splits exceed the grid, so the splits of a tile are not co-resident):
// out[m, n] = sum_z partial[z][m][n] (fixed order: bit-stable) + bias + rowvec + residual -> bf16
__device__ __noinline__ void splitk_reduce_op(const SeqEW& p) {
  const float* partial = reinterpret_cast<const float*>(p.src);
  const bf16* residual = reinterpret_cast<const bf16*>(p.src2);
  bf16* out = reinterpret_cast<bf16*>(p.dst);
  const int splits = p.aux, M = p.n_img, N = p.c, hw = p.h, rowvec_ld = p.aux2, res_ld = p.aux3, ldo = p.aux4;
  const int nv = N / 8;
  const int64_t total = (int64_t)M * nv;
  for (int64_t i = (int64_t)blockIdx.x * SEQ_THREADS + threadIdx.x; i < total; i += (int64_t)gridDim.x * SEQ_THREADS) {
    const int m = (int)(i / nv), n = (int)(i % nv) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int z = 0; z < splits; ++z) {
      const float4* pp = reinterpret_cast<const float4*>(partial + ((int64_t)z * M + m) * N + n);
      const float4 a = __ldcg(pp), b = __ldcg(pp + 1);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (p.bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += p.bias[n + j];
    }
    if (p.rowvec) {
      const float* rv = p.rowvec + (int64_t)(m / hw) * rowvec_ld + n;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += __ldcg(rv + j);
    }
    if (residual) {
      float f[8];
      unpack8(ld_cg16(residual + (int64_t)m * res_ld + n), f);
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += f[j];
    }
    *reinterpret_cast<uint4*>(out + (int64_t)m * ldo + n) = pack8(v);
  }
}

// =====================================================================================================
// the kernel
// =====================================================================================================
__global__ void __launch_bounds__(SEQ_THREADS, 1) seq_kernel(const SeqOp* __restrict__ ops, int n_ops, unsigned* sync,
                                                             long long* timing) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = tc::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      tc::mbar_init(smem_base + OFF_BAR_FULL + 8 * s, 1);
      tc::mbar_init(smem_base + OFF_BAR_EMPTY + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(smem_base + OFF_ACC_FULL + 8 * b, 1);
      tc::mbar_init(smem_base + OFF_ACC_EMPTY + 8 * b, 128);
    }
    tc::mbar_fence_init();
    if (timing != nullptr && blockIdx.x == 0) stamp(timing, 0);
  }
  if (warp == 1) tc::tmem_alloc<512>(smem_base + OFF_TMEM);
  if (warp == 2) reinterpret_cast<uint4*>(smem + OFF_DESC)[lane] = __ldg(reinterpret_cast<const uint4*>(ops) + lane);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem + OFF_TMEM);
  const SeqOpC& op = *reinterpret_cast<const SeqOpC*>(smem + OFF_DESC);

  if (threadIdx.x == 0 && op.type == SEQ_GEMM) acquire_tensormaps(ops, op.g.nseg);
  GridBarrier gb{sync, 0u, timing};
  RoleState st{0xffffffffu, 0u, 0u};
  for (int i = 0; i < n_ops; ++i) {
    const int type = op.type;
    if (type == SEQ_GEMM) {
      gemm_op(op.g, op.pf, op.next_nseg, ops + i, smem, smem_base, tmem_base, st);
    } else {
      if (threadIdx.x == 0) {
        acquire_tensormaps(ops + i + 1, op.next_nseg);
        prefetch_next_weights(op.pf);
      }
      if (type == SEQ_GN) {
        if (op.gn.warp_mode) gn_warp(op.gn);
        else gn_cta(op.gn, smem, gb);
      } else if (type == SEQ_LN) {
        const int c = op.ln.c;
        if (c <= 512) ln_rows<2>(op.ln);
        else if (c <= 768) ln_rows<3>(op.ln);
        else if (c <= 1280) ln_rows<5>(op.ln);
        else ln_rows<8>(op.ln);
      } else if (type == SEQ_UPSAMPLE) {
        upsample_op(op.ew);
      } else if (type == SEQ_IM2COL) {
        im2col_op(op.ew);
      } else if (type == SEQ_SINUSOID) {
        sinusoid_op(op.ew);
      } else if (type == SEQ_SPLITK_REDUCE) {
        splitk_reduce_op(op.ew);
      }
    }
    if (i + 1 < n_ops) gb.sync(ops + i + 1, smem + OFF_DESC);
  }
  // ---- teardown: the last CTA out re-arms the barrier words for the next launch (every CTA that got here has passed
  // every barrier, so nobody can still be polling)
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
  }
  if (threadIdx.x == 0) {
    if (timing != nullptr && blockIdx.x == 0) stamp(timing, gb.n + 1);  // [0] start, [k] after barrier k, then end
    __threadfence();
    if (atomicAdd(&sync[1], 1u) == gridDim.x - 1) {
      sync[0] = 0u;
      sync[1] = 0u;
      __threadfence();
    }
  }
}

}  // namespace

// =====================================================================================================
// host: planning and launch
// =====================================================================================================
int seq_grid() {
  static int sms[kMaxDevices] = {};
  int dev = 0;
  MV_CUDA(cudaGetDevice(&dev));
  MV_CHECK(dev >= 0 && dev < kMaxDevices, "device index out of range");
  if (sms[dev] == 0) MV_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
  return sms[dev];
}

namespace {

// ---- tile / split-K selection ---------------------------------------------------------------------
// Measured on B200 (profiles/r01_*): one SM ingests at most ~150 GB/s of TMA traffic from L2 and ~4.3 TMA boxes/us, so a
// GEMM here is bound by (bytes the busiest SM has to pull) long before the tensor pipe saturates.  The model below picks
// the N-tile (arithmetic intensity per SM) and the split-K factor (SMs kept busy) that minimise
//   max(load time, MMA time) + split-K reduction time.
struct TileChoice {
  int bn, splits;
};

int count_steps(const mvldm_gemm_desc& d) {
  int steps = 0;
  for (int i = 0; i < d.nseg; ++i) steps += d.seg[i].ntaps * ceil_div(d.seg[i].c / BK, KC);
  return steps;
}

TileChoice pick_tiles(const mvldm_gemm_desc& d) {
  static const int kBN[5] = {256, 160, 128, 64, 32};
  static const int kSplits[12] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32};
  const int M = d.n_img * d.oh * d.ow, mt = ceil_div(M, BM), num_steps = count_steps(d);
  const double grid = (double)seq_grid();
  const double kb_per_step = (double)(d.k / BK) / num_steps;  // 64-chunks an average step carries (<= KC)
  TileChoice best{0, 1};
  double best_t = 1e30;
  // tools/gemm_sweep.py: force one configuration to measure it against the model's choice
  const char* force_bn = getenv("MVLDM_GEMM_BN");
  const char* force_sp = getenv("MVLDM_GEMM_SPLITS");
  for (int bn : kBN) {
    if (d.n % bn != 0) continue;
    if (d.mode == 2 && bn != 32) continue;
    if (force_bn && atoi(force_bn) != bn) continue;
    for (int sp : kSplits) {
      if (sp > 1 && (d.mode != 0 || num_steps / sp < 3)) break;
      if (force_sp && atoi(force_sp) != sp) continue;
      const int st_per = ceil_div(num_steps, sp), splits = ceil_div(num_steps, st_per);
      const double ctas = (double)mt * (d.n / bn) * splits;
      const double per_sm = std::ceil(ctas / grid);  // work items the busiest SM runs
      const double step_bytes = kb_per_step * (A_BYTES + bn * 128.0);
      const double t_step = std::max(std::max(2.0 / 4.3e6, step_bytes / 150e9), kb_per_step * 4.0 * (bn / 2.0) / 1.9e9);
      const double main = st_per * t_step;
      const double epi = bn * (d.mode == 1 ? 12e-9 : 6e-9);  // TMEM -> registers -> global, per item
      // persistent CTA: ramp once, items back to back (epilogue hidden behind the next main loop), last epilogue exposed
      const double t_sm = 1.5e-6 + per_sm * std::max(main, epi) + epi;
      const double t_red = splits > 1 ? (splits + 1.0) * M * (double)d.n * 4.0 / 3e12 + 3e-6 : 0.0;
      const double t = t_sm + t_red;
      if (t < best_t) {
        best_t = t;
        best = TileChoice{bn, splits};
      }
    }
  }
  return best;
}

constexpr size_t kCounterBytes = 2 * 4096 * sizeof(int);  // arrive/done counters live at the head of the workspace

}  // namespace

size_t seq_gemm_workspace_bytes(const mvldm_gemm_desc& d) {
  if (count_steps(d) == 0) return 0;
  const int splits = pick_tiles(d).splits;
  return splits > 1 ? kCounterBytes + (size_t)splits * d.n_img * d.oh * d.ow * d.n * sizeof(float) : 0;
}

bool seq_plan_gemm(const mvldm_gemm_desc& d, void* workspace, size_t workspace_bytes, SeqOp& op, SeqOp& reduce) {
  memset(&op, 0, sizeof(op));
  op.c.type = SEQ_GEMM;
  SeqGemm& p = op.c.g;
  const int hw = d.oh * d.ow;
  p.M = d.n_img * hw;
  p.N = d.n;
  p.hw = hw;
  p.ow = d.ow;
  MV_CHECK(d.nseg >= 1 && d.nseg <= MVLDM_MAX_SEGS, "gemm: bad segment count");
  MV_CHECK(d.ow <= BM && BM % d.ow == 0, "gemm: output width must divide 128");
  MV_CHECK(hw % BM == 0 || BM % hw == 0, "gemm: pixels per image must divide or be a multiple of 128");
  // 128-pixel tile = bw x bh x bn box of whole rows / whole images
  const int bw = d.ow;
  const int bh = std::min(d.oh, BM / bw);
  const int bimg = BM / (bw * bh);
  for (int i = 0; i < d.nseg; ++i)
    MV_CHECK(d.seg[i].c >= BK && d.seg[i].c % BK == 0, "gemm: segment channels must be a multiple of 64");
  const TileChoice tile = pick_tiles(d);
  const int BN = tile.bn;
  MV_CHECK(BN != 0, "gemm: N must be a multiple of 32");
  int ktot = 0;
  for (int i = 0; i < d.nseg; ++i) {
    const mvldm_aseg& a = d.seg[i];
    MV_CHECK(a.c % BK == 0 && a.ctot % BK == 0, "gemm: segment channels must be a multiple of 64");
    MV_CHECK(a.stride == 1 || a.stride == 2, "gemm: stride must be 1 or 2");
    MV_CHECK(a.sh == d.oh * a.stride && a.sw == d.ow * a.stride, "gemm: source / output size mismatch");
    MV_CHECK((reinterpret_cast<uintptr_t>(a.ptr) & 15) == 0, "gemm: source pointer must be 16-byte aligned");
    MV_CHECK(a.ntaps >= 1 && a.ntaps <= 9, "gemm: 1..9 taps per segment");
    SeqSeg& t = p.seg[i];
    t.ncblk = (int16_t)(a.c / BK);
    t.ntaps = (int16_t)a.ntaps;
    t.stride = (int16_t)a.stride;
    t.spt = (int16_t)ceil_div(a.c / BK, KC);
    t.kchunk0 = (int16_t)(ktot / BK);
    for (int j = 0; j < a.ntaps; ++j) {
      t.dh[j] = a.dh[j];
      t.dw[j] = a.dw[j];
      MV_CHECK(a.coff[j] % BK == 0, "gemm: tap channel offset must be a multiple of 64");
      t.cblk[j] = (int16_t)(a.coff[j] / BK);
    }
    // 5-D view (64 channels, w, h, image, 64-channel chunk): the chunk is the slowest box dimension so that one
    // box lands as [chunk][pixel][128 B] = consecutive K-major SW128 operand tiles
    const uint64_t dims[5] = {(uint64_t)BK, (uint64_t)a.sw, (uint64_t)a.sh, (uint64_t)d.n_img, (uint64_t)(a.ctot / BK)};
    const uint64_t strides[4] = {(uint64_t)a.ctot * 2, (uint64_t)a.sw * a.ctot * 2, (uint64_t)a.sh * a.sw * a.ctot * 2,
                                 (uint64_t)BK * 2};
    const uint32_t es[5] = {1, (uint32_t)a.stride, (uint32_t)a.stride, 1, 1};
    for (int kc = 1; kc <= KC; ++kc) {
      const uint32_t box[5] = {(uint32_t)BK, (uint32_t)(bw * a.stride), (uint32_t)(bh * a.stride), (uint32_t)bimg, (uint32_t)kc};
      op.tmA[i][kc - 1] = make_tmap_bf16(a.ptr, 5, dims, strides, box, es);
    }
    ktot += a.c * a.ntaps;
  }
  p.nseg = d.nseg;
  MV_CHECK(ktot == d.k, "gemm: K mismatch between segments and weights");
  MV_CHECK(d.k / BK < 65536, "gemm: K too large");
  p.num_steps = count_steps(d);
  MV_CHECK(d.mode != 2 || d.n == 32, "gemm: NCHW head output expects N padded to 32");
  {
    const uint64_t dims[3] = {(uint64_t)BK, (uint64_t)d.n, (uint64_t)(d.k / BK)};
    const uint64_t strides[2] = {(uint64_t)d.k * 2, (uint64_t)BK * 2};
    const uint32_t es[3] = {1, 1, 1};
    for (int kc = 1; kc <= KC; ++kc) {
      const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, (uint32_t)kc};
      op.tmB[kc - 1] = make_tmap_bf16(d.w, 3, dims, strides, box, es);
    }
  }
  int splits = tile.splits;
  if (splits > 1 && seq_gemm_workspace_bytes(d) > workspace_bytes) splits = 1;  // no scratch: plain single-pass GEMM
  p.steps_per_split = ceil_div(p.num_steps, splits);
  splits = ceil_div(p.num_steps, p.steps_per_split);
  MV_CHECK(splits <= SEQ_MAX_SPLITS, "gemm: too many K splits");
  p.mt = ceil_div(p.M, BM);
  p.nt = d.n / BN;
  p.splits = splits;
  p.bn = BN;
  p.stages = stages_for(BN);
  p.partial = splits > 1 ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes) : nullptr;
  // the fused reduction needs every split of a tile resident at once: one work item per CTA of the co-resident grid
  const int work = p.mt * p.nt * splits;
  const bool fused = splits > 1 && work <= seq_grid() && p.mt * p.nt <= 4096;
  p.counters = fused ? reinterpret_cast<int*>(workspace) : nullptr;
  p.bias = d.bias;
  p.rowvec = d.rowvec;
  p.rowvec_ld = d.rowvec_ld;
  p.residual = reinterpret_cast<const bf16*>(d.residual);
  p.res_ld = d.res_ld;
  p.mode = d.mode;
  p.out = d.out;
  p.ldo = d.ldo;
  p.n_valid = d.n_valid;
  MV_CHECK(d.mode >= 0 && d.mode <= 4, "gemm: bad output mode");
  if (d.mode == 0 || d.mode == 3 || d.mode == 4)
    MV_CHECK(d.ldo % 8 == 0 && (!d.residual || d.res_ld % 8 == 0), "gemm: row pitch must be a multiple of 8");
  if (splits > 1 && !fused) {
    memset(&reduce, 0, sizeof(reduce));
    reduce.c.type = SEQ_SPLITK_REDUCE;
    SeqEW& r = reduce.c.ew;
    r.src = p.partial;
    r.src2 = p.residual;
    r.dst = p.out;
    r.bias = p.bias;
    r.rowvec = p.rowvec;
    r.n_img = p.M; r.c = p.N; r.h = p.hw;
    r.aux = splits; r.aux2 = p.rowvec_ld; r.aux3 = p.res_ld; r.aux4 = p.ldo;
    return true;
  }
  return false;
}

void seq_link_prefetch(SeqOp* ops, int n, const SeqOp* dev_ops) {
  for (int i = 0; i + 1 < n; ++i) {
    SeqOp& cur = ops[i];
    const SeqOp& nx = ops[i + 1];
    cur.c.next_nseg = 0;
    cur.c.pf.map = nullptr;
    if (nx.c.type != SEQ_GEMM) continue;
    const SeqGemm& g = nx.c.g;
    cur.c.next_nseg = g.nseg;
    SeqPrefetch& pf = cur.c.pf;
    pf.map = &dev_ops[i + 1].tmB[KC - 1];
    pf.bn = g.bn; pf.mt = g.mt; pf.nt = g.nt; pf.splits = g.splits;
    // K chunk (64 wide) at which every pipeline step starts: chunks are consecutive along K over (segment, tap, block)
    std::vector<int> step_chunk;
    for (int s = 0; s < g.nseg; ++s)
      for (int t = 0; t < g.seg[s].ntaps; ++t)
        for (int cb = 0; cb < g.seg[s].ncblk; cb += KC) step_chunk.push_back(g.seg[s].kchunk0 + t * g.seg[s].ncblk + cb);
    const int total = g.seg[g.nseg - 1].kchunk0 + g.seg[g.nseg - 1].ntaps * g.seg[g.nseg - 1].ncblk;
    for (int z = 0; z < g.splits; ++z) pf.chunk0[z] = (uint16_t)step_chunk[std::min<size_t>((size_t)z * g.steps_per_split, step_chunk.size() - 1)];
    pf.chunk0[g.splits] = (uint16_t)total;
  }
  if (n > 0) {
    ops[n - 1].c.next_nseg = 0;
    ops[n - 1].c.pf.map = nullptr;
  }
  for (int i = 0; i < n; ++i) ops[i].c.index = i;
}

size_t seq_groupnorm_scratch_floats(int n_img, int c, int groups) {
  (void)c; (void)groups;
  return (size_t)n_img * 4096;  // >= 8 floats per (image, channel block, pixel chunk) item for every supported shape
}

void seq_plan_groupnorm(const bf16* x0, int c0, const bf16* x1, int c1, int n_img, int hw, int groups, float eps,
                        const float* gamma, const float* beta, bool silu, bf16* out, float* scratch, int grid, SeqOp& op) {
  memset(&op, 0, sizeof(op));
  op.c.type = SEQ_GN;
  SeqGN& p = op.c.gn;
  const int C = c0 + c1;
  MV_CHECK(groups > 0 && C % groups == 0, "groupnorm: channels not divisible by groups");
  MV_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && c0 > 0, "groupnorm: channel counts must be multiples of 8");
  const int cgn = C / groups;
  MV_CHECK(cgn >= 8, "groupnorm: groups narrower than 8 channels are not supported");
  int cb = cgn;
  while (cb % 8 != 0) cb += cgn;  // lcm(group width, 8): whole groups and whole 16-byte vectors
  MV_CHECK(cb / cgn <= 4 && C % cb == 0 && cb / 8 <= 32, "groupnorm: unsupported group width (need lcm(C/groups, 8) <= 4 groups)");
  const int NV = cb / 8, nblk = C / cb;
  p.x0 = x0; p.x1 = x1; p.gamma = gamma; p.beta = beta; p.out = out; p.partial = scratch;
  p.c0 = c0; p.c1 = c1; p.n_img = n_img; p.hw = hw; p.cgn = cgn; p.cb = cb;
  p.silu = silu ? 1 : 0;
  p.eps = eps;
  p.nv_magic = (uint32_t)((0x100000000ull + NV - 1) / NV);
  const int64_t items0 = (int64_t)n_img * nblk;
  if (hw * NV <= 32 * GN_WARP_MAXV) {
    p.warp_mode = 1;
    p.ps = 1;
    p.px = hw;
    p.n_items = (int)items0;
    return;
  }
  int ps = 1;
  while (items0 * ps * 2 <= grid && hw % (ps * 2) == 0 && hw / (ps * 2) >= 64) ps *= 2;
  p.ps = ps;
  p.px = hw / ps;
  p.n_items = (int)(items0 * ps);
  MV_CHECK((size_t)p.n_items * 8 <= seq_groupnorm_scratch_floats(n_img, C, groups) || ps == 1, "groupnorm: scratch too small");
  const int64_t per_cta = (p.n_items + grid - 1) / grid;
  p.cache = per_cta * p.px * cb * 2 <= WORK_BYTES ? 1 : 0;
}

void seq_plan_layernorm(const bf16* x, int rows, int c, float eps, const float* gamma, const float* beta, bf16* out, SeqOp& op) {
  MV_CHECK(c % 8 == 0 && c <= 32 * 8 * 8, "layernorm: unsupported channel count");
  memset(&op, 0, sizeof(op));
  op.c.type = SEQ_LN;
  SeqLN& p = op.c.ln;
  p.x = x; p.gamma = gamma; p.beta = beta; p.out = out; p.rows = rows; p.c = c; p.eps = eps;
}

void seq_plan_upsample(const bf16* x, int n_img, int h, int w, int c, bf16* out, SeqOp& op) {
  MV_CHECK(c % 8 == 0, "upsample: channels must be a multiple of 8");
  memset(&op, 0, sizeof(op));
  op.c.type = SEQ_UPSAMPLE;
  SeqEW& p = op.c.ew;
  p.src = x; p.dst = out; p.n_img = n_img; p.h = h; p.w = w; p.c = c;
}

void seq_plan_im2col(const float* latents, int n_img, int cin, int h, int w, int kpad, bf16* out, SeqOp& op) {
  MV_CHECK(kpad >= 9 * cin, "im2col: kpad too small");
  MV_CHECK(kpad % 8 == 0 && (int64_t)n_img * h * w * kpad < (1ll << 31), "im2col: kpad must be a multiple of 8 (32-bit indexing)");
  memset(&op, 0, sizeof(op));
  op.c.type = SEQ_IM2COL;
  SeqEW& p = op.c.ew;
  p.src = latents; p.dst = out; p.n_img = n_img; p.h = h; p.w = w; p.c = cin; p.aux = kpad;
}

void seq_plan_sinusoid(const int64_t* t, int n, int dim, bf16* out, SeqOp& op) {
  memset(&op, 0, sizeof(op));
  op.c.type = SEQ_SINUSOID;
  SeqEW& p = op.c.ew;
  p.src = t; p.dst = out; p.n_img = n; p.c = dim;
}

void seq_configure() {  // once per device, outside any stream capture
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    MV_CUDA(cudaFuncSetAttribute(seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int per_sm = 0;
    MV_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, seq_kernel, SEQ_THREADS, SMEM_BYTES));
    MV_CHECK(per_sm >= 1, "seq_launch: the sequence kernel does not fit on an SM of this device");
  }
}

void seq_launch(cudaStream_t s, const SeqOp* dev_ops, int n_ops, unsigned* sync, long long* timing) {
  MV_CHECK(n_ops > 0, "seq_launch: empty op list");
  seq_configure();
  // cooperative launch: the grid barriers need every CTA resident; the runtime refuses the launch otherwise
  static const bool coop = [] {
    const char* e = getenv("MVLDM_SEQ_COOP");
    return !e || atoi(e) != 0;
  }();
  void* args[] = {(void*)&dev_ops, (void*)&n_ops, (void*)&sync, (void*)&timing};
  if (coop) {
    MV_CUDA(cudaLaunchCooperativeKernel((const void*)seq_kernel, dim3(seq_grid()), dim3(SEQ_THREADS), args, SMEM_BYTES, s));
  } else {
    MV_CUDA(cudaLaunchKernel((const void*)seq_kernel, dim3(seq_grid()), dim3(SEQ_THREADS), args, SMEM_BYTES, s));
  }
  MV_LAUNCHED();
}

void seq_run_host_ops(cudaStream_t s, const SeqOp* host_ops, int n_ops) {
  struct Slot {
    SeqOp* ops = nullptr;
    int cap = 0;
    unsigned* sync = nullptr;
  };
  static Slot slots[kMaxDevices];
  int dev = 0;
  MV_CUDA(cudaGetDevice(&dev));
  Slot& sl = slots[dev];
  if (!sl.sync) {
    MV_CUDA(cudaMalloc(&sl.sync, 2 * sizeof(unsigned)));
    MV_CUDA(cudaMemset(sl.sync, 0, 2 * sizeof(unsigned)));
  }
  if (n_ops > sl.cap) {
    MV_CUDA(cudaDeviceSynchronize());
    if (sl.ops) cudaFree(sl.ops);
    sl.cap = std::max(n_ops, 8);
    MV_CUDA(cudaMalloc(&sl.ops, sizeof(SeqOp) * sl.cap));
  }
  std::vector<SeqOp> tmp(host_ops, host_ops + n_ops);
  seq_link_prefetch(tmp.data(), n_ops, sl.ops);
  MV_CUDA(cudaMemcpyAsync(sl.ops, tmp.data(), sizeof(SeqOp) * n_ops, cudaMemcpyHostToDevice, s));
  seq_launch(s, sl.ops, n_ops, sl.sync, nullptr);
}

// ---- op-level entry points (C ABI mvldm_op_*; the tests drive the same kernel the forward uses) ----------------
size_t gemm_tc_workspace_bytes(const mvldm_gemm_desc& d) { return seq_gemm_workspace_bytes(d); }

void gemm_tc(cudaStream_t s, const mvldm_gemm_desc& d, void* workspace, size_t workspace_bytes) {
  SeqOp ops[2];
  const bool red = seq_plan_gemm(d, workspace, workspace_bytes, ops[0], ops[1]);
  seq_run_host_ops(s, ops, red ? 2 : 1);
}

size_t groupnorm_scratch_floats(int n_img, int groups) { return seq_groupnorm_scratch_floats(n_img, 0, groups); }

void groupnorm(cudaStream_t s, const bf16* x0, int c0, const bf16* x1, int c1, int n_img, int hw, int groups, float eps,
               const float* gamma, const float* beta, bool silu, bf16* out, float* scratch) {
  SeqOp op;
  seq_plan_groupnorm(x0, c0, x1, c1, n_img, hw, groups, eps, gamma, beta, silu, out, scratch, seq_grid(), op);
  seq_run_host_ops(s, &op, 1);
}

void layernorm(cudaStream_t s, const bf16* x, int rows, int c, float eps, const float* gamma, const float* beta, bf16* out) {
  SeqOp op;
  seq_plan_layernorm(x, rows, c, eps, gamma, beta, out, op);
  seq_run_host_ops(s, &op, 1);
}

}  // namespace mvldm
