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
// kernel behaviour switches (MVLDM_SEQ_FLAGS, default all on): measured one by one in profiles/
constexpr int SEQ_F_TMAP_FENCE = 1, SEQ_F_PROXY_FENCE = 2, SEQ_F_PREFETCH = 4;

// ---- shared memory map (dynamic, carved by hand; base rounded up to 1024 B for the SWIZZLE_128B atoms) ----
//   [0, AUX)                     op descriptor, mbarriers, TMEM slot, reduction scratch
//   [AUX, AUX + WORK)            GEMM: `stages` ring slots of KC*(A tile + B tile), then the bias/row-vector table;
//                                GroupNorm: the CTA's channel slabs
constexpr int OFF_DESC = 0;
constexpr int OFF_BAR_FULL = 512, OFF_BAR_EMPTY = 576, OFF_ACC_FULL = 640, OFF_ACC_EMPTY = 656, OFF_TMEM = 672;
constexpr int OFF_RED = 704;    // [8 warps][4] floats
constexpr int OFF_STAT = 832;   // [8 item groups][4][2] floats
constexpr int AUX_BYTES = 2048;
constexpr int WORK_BYTES = 3 * KC * (A_BYTES + 160 * 128) + CV_IMGS * 160 * 4;  // 3 stages of the 160-wide tile + its table
constexpr int SMEM_BYTES = 1024 + AUX_BYTES + WORK_BYTES;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA may use");

inline int stage_bytes_for(int bn) { return KC * (A_BYTES + bn * 128); }
inline int stages_for(int bn) { return std::min(MAX_STAGES, (WORK_BYTES - CV_IMGS * bn * 4) / stage_bytes_for(bn)); }

// ---- small device helpers ---------------------------------------------------------------------------------
// CODE SIZE IS A FIRST-ORDER COST HERE.  Every op body runs once per launch position, usually on a cold instruction
// cache (the kernel is far larger than the 32 KB L1.5 I-cache, and attention kernels run in between): measured on B200,
// cold straight-line code retires at ~9 cycles per instruction, so a 40 KB unrolled op body cost ~8 us by itself.
// Hence: rolled loops over shared memory (cp.async staging gives the memory-level parallelism that unrolling would),
// 16-column epilogue chunks, one code path per op.  Check `cuobjdump -elf build/seq.o` symbol sizes after every change.
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// erf-form GELU (F.gelu default, mvdream/attention.py:60-70) with erf from Abramowitz & Stegun 7.1.26 (|abs err| <
// 1.5e-7, far below the bf16 the result is stored in): Phi(-|x|) = 0.5 * poly(t) * exp(-x^2/2), t = 1/(1 + p |x|/sqrt2).
// 18 instructions (2 MUFU) instead of erff's ~35 with branches.
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
// 16-byte global -> shared copy that bypasses L1 and registers: issued from a rolled loop, all copies stay in flight
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc::smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// 256-bit global store: one instruction covers a full 32-byte sector per thread (rows are >= 64 B apart, so 16-byte
// stores would touch every sector twice)
__device__ __forceinline__ void st_global_v8(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                             uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}
// 32 lanes x 8 columns (registers <-> TMEM): the residual row is parked in the accumulator buffer's spare columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(a.x), "r"(a.y),
               "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}

// tcgen05.wait::ld that the compiler cannot move register reads across: the TMEM load writes its destination registers
// asynchronously, so every use of them must stay behind the wait - which needs a data dependency, not just asm volatile
// (an fp32 accumulator chunk read as a plain bit cast was hoisted above the wait and picked up in-flight registers)
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])::"memory");
}
__device__ __forceinline__ void tmem_ld_dep16(uint32_t (&r)[16]) {  // after the wait: ties a second destination array to it
  asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
               "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])::"memory");
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
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
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
// Measured (tools/seq_barrier_bench.py, tools/seq_trace.py): ~1 us from the last CTA's arrival to the release.
// All arguments by value (registers): a struct passed by reference would live in local memory, which is L2 here.
// Returns the new barrier count.
__device__ __noinline__ unsigned grid_sync(unsigned* ctr, unsigned n, int flags, long long* timing, long long* trace, int op,
                                           const SeqOp* next, uint8_t* desc_smem) {
  long long* tr = trace && threadIdx.x == 0 ? trace + ((size_t)n * gridDim.x + blockIdx.x) * 16 : nullptr;
  if (tr) tr[0] = globaltimer_ns();
  // generic-proxy accesses of this op (global stores other CTAs will read with TMA; shared-memory staging tiles, slabs and
  // row buffers that the next op's TMA loads will overwrite) are ordered before the async-proxy accesses that follow
  if (flags & 8) __threadfence();  // debug: every thread fences (the pre-rewrite barrier)
  if (flags & SEQ_F_PROXY_FENCE) asm volatile("fence.proxy.async;" ::: "memory");
  __syncthreads();
  if (tr) tr[1] = globaltimer_ns();
  ++n;
  if (threadIdx.x == 0) {
    // cumulative gpu-scope fence: publishes the global writes of the whole CTA (ordered before it by the CTA barrier)
    asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
  }
  if (next != nullptr && threadIdx.x >= 32 && threadIdx.x < 32 + SEQ_OPC_BYTES / 16)
    reinterpret_cast<uint4*>(desc_smem)[threadIdx.x - 32] = __ldg(reinterpret_cast<const uint4*>(next) + (threadIdx.x - 32));
  if (threadIdx.x == 0) {
    const unsigned target = n * gridDim.x;
    unsigned spins = 0;
    while (ld_acquire_u32(ctr) < target) {
      if (++spins > (1u << 25)) __trap();  // a protocol bug / lost co-residency traps instead of hanging the GPU
    }
    if (timing != nullptr && blockIdx.x == 0) stamp(timing, n);
    if (tr) {
      tr[2] = globaltimer_ns();
      tr[3] = op;
      tr[5] = spins;
    }
  }
  __syncthreads();
  if (flags & SEQ_F_PROXY_FENCE) asm volatile("fence.proxy.async;" ::: "memory");
  return n;
}

struct GridBarrier {  // the barrier words of a launch, carried by value
  unsigned* ctr;
  unsigned n;  // barriers passed
  long long* timing;
  int flags;
  long long* trace;  // debug: [barrier][cta][16] stamps (tools/seq_trace.py)
  int op;
};

// role state that survives from op to op: mbarrier phase parities per ring slot, accumulator-buffer counter
struct RoleState {
  uint32_t empty_par;  // producer: parity to wait for on bar_empty[s] before refilling slot s
  uint32_t full_par;   // MMA issuer: parity to wait for on bar_full[s]
  long long* tr;       // debug timeline row of the barrier that will close this op (slots 4..15: clock64 since t0), or NULL
  long long t0;        // clock64 when this op started (debug)
};

// descriptor-cache prefetch of every tensor map op `o` uses (a cold map costs the first TMA ~2 us: the 128-byte
// descriptors are touched once per forward, so they come from HBM)
__device__ __forceinline__ void prefetch_tensormaps(const SeqOp* o, int nseg) {
#pragma unroll 1
  for (int i = 0; i < (nseg + 1) * KC; ++i) {
    const CUtensorMap* m = i < nseg * KC ? &o->tmA[0][0] + i : &o->tmB[0] + (i - nseg * KC);
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
  }
}

// L2 prefetch of the next GEMM op's weight tiles: its [bn x KC chunk] boxes are dealt round-robin to the CTAs, so every
// SM's TMA unit carries an equal share (a box costs TMA issue time whether or not the data returns to the SM)
__device__ __forceinline__ void prefetch_next_weights(const SeqPrefetch& pf) {
  if (pf.map == nullptr) return;
  const int per_n = (pf.nchunks + KC - 1) / KC;  // boxes along K for one n-tile
  const int boxes = pf.nt * per_n;
#pragma unroll 1
  for (int b = blockIdx.x; b < boxes; b += gridDim.x) {
    const int ntile = b / per_n, kb = b - ntile * per_n;
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(pf.map)),
                 "r"(0), "r"(ntile * pf.bn), "r"(kb * KC)
                 : "memory");
  }
}

// =====================================================================================================
// SEQ_GEMM
// =====================================================================================================
// (segment, tap, channel block) walk shared by the producer and the MMA issuer
struct KWalk {
  int s, t, cb;
  __device__ __forceinline__ void start(const SeqGemm& g, int step) {
    s = 0;
    cb = step;
    while (cb >= g.seg[s].ntaps * g.seg[s].spt) {
      cb -= g.seg[s].ntaps * g.seg[s].spt;
      ++s;
    }
    t = cb / g.seg[s].spt;
    cb = (cb - t * g.seg[s].spt) * KC;
  }
  __device__ __forceinline__ int chunks(const SeqGemm& g) const { return min(KC, (int)g.seg[s].ncblk - cb); }
  __device__ __forceinline__ void next(const SeqGemm& g, int kc) {
    cb += kc;
    if (cb == g.seg[s].ncblk) {
      cb = 0;
      if (++t == g.seg[s].ntaps) {
        t = 0;
        ++s;
      }
    }
  }
};

__device__ __noinline__ void gemm_producer(const SeqGemm& g, const SeqOp* opg, uint32_t smem_base, RoleState& st) {
  const int BN = g.bn;
  const uint32_t b_bytes = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = KC * (A_BYTES + b_bytes);
  const uint32_t ring_base = smem_base + AUX_BYTES;
  const uint32_t bar_full = smem_base + OFF_BAR_FULL, bar_empty = smem_base + OFF_BAR_EMPTY;
  const int num_work = g.mt * g.nt * g.splits;
  // role state and debug pointers live in registers inside this function: `st` sits in local memory (= L2, the L1 is
  // carved out for shared memory) and every asm with a memory clobber would force a reload
  uint32_t empty_par = st.empty_par;
  long long* const tr = st.tr;
  const long long t0 = st.t0;
  const int stages = g.stages;
  int ring = 0;
#pragma unroll 1
  for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
    const int mtile = w % g.mt, ntile = (w / g.mt) % g.nt, z = w / (g.mt * g.nt);
    const int m0 = mtile * BM, n0 = ntile * BN;
    const int st_begin = z * g.steps_per_split;
    const int nst = min(g.num_steps - st_begin, g.steps_per_split);
    const int img0 = m0 / g.hw;
    const int y0 = (m0 - img0 * g.hw) / g.ow;
    KWalk k;
    k.start(g, st_begin);
#pragma unroll 1
    for (int i = 0; i < nst; ++i) {
      const SeqSeg& sg = g.seg[k.s];
      const int kc = k.chunks(g);
      tc::mbar_wait(bar_empty + 8 * ring, (empty_par >> ring) & 1u);
      empty_par ^= 1u << ring;
      const uint32_t full = bar_full + 8 * ring;
      tc::mbar_expect_tx(full, kc * (A_BYTES + b_bytes));
      const uint32_t sa = ring_base + ring * stage_bytes;
      tc::tma_load_5d(sa, &opg->tmA[k.s][kc - 1], full, 0, sg.dw[k.t], y0 * sg.stride + sg.dh[k.t], img0, sg.cblk[k.t] + k.cb);
      tc::tma_load_3d(sa + KC * A_BYTES, &opg->tmB[kc - 1], full, 0, n0, sg.kchunk0 + k.t * sg.ncblk + k.cb);
      if (tr && i == 0 && w == (int)blockIdx.x) tr[6] = clock64() - t0;  // first loads issued
      ring = ring + 1 == stages ? 0 : ring + 1;
      k.next(g, kc);
    }
  }
  st.empty_par = empty_par;
}

// the whole warp runs the warp-uniform loop and the barrier waits; one elected lane issues tcgen05.mma / commit,
// which lets ptxas emit the UTCHMMAs back to back instead of one ELECT/branch loop per instruction
__device__ __noinline__ void gemm_mma(const SeqGemm& g, uint32_t smem_base, uint32_t tmem_base, RoleState& st, uint32_t acc_n) {
  const int BN = g.bn;
  const uint32_t b_bytes = (uint32_t)BN * 128u;
  const uint32_t stage_bytes = KC * (A_BYTES + b_bytes);
  const uint32_t ring_base = smem_base + AUX_BYTES;
  const uint32_t bar_full = smem_base + OFF_BAR_FULL, bar_empty = smem_base + OFF_BAR_EMPTY;
  const uint32_t bar_acc_full = smem_base + OFF_ACC_FULL, bar_acc_empty = smem_base + OFF_ACC_EMPTY;
  const int num_work = g.mt * g.nt * g.splits;
  const uint32_t idesc = tc::umma_idesc_bf16(BM, BN, false, false);
  uint32_t full_par = st.full_par;  // registers, see gemm_producer
  long long* const tr = st.tr;
  const long long t0 = st.t0;
  const int stages = g.stages;
  int ring = 0;
#pragma unroll 1
  for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
    const int z = w / (g.mt * g.nt);
    const int st_begin = z * g.steps_per_split;
    const int nst = min(g.num_steps - st_begin, g.steps_per_split);
    KWalk k;  // same walk as the producer, to know how many chunks each step carries
    k.start(g, st_begin);
    const uint32_t ab = acc_n & 1u;
    tc::mbar_wait(bar_acc_empty + 8 * ab, ((acc_n >> 1) & 1u) ^ 1u);  // epilogue has drained this buffer
    tc::tc_fence_after();
    const uint32_t tmem_d = tmem_base + ab * ACC_COLS;
#pragma unroll 1
    for (int i = 0; i < nst; ++i) {
      const int kc = k.chunks(g);
      tc::mbar_wait(bar_full + 8 * ring, (full_par >> ring) & 1u);
      full_par ^= 1u << ring;
      tc::tc_fence_after();
      if (tr && i == 0 && w == (int)blockIdx.x && (threadIdx.x & 31) == 0) tr[7] = clock64() - t0;  // first tile landed
      const uint32_t sa = ring_base + ring * stage_bytes;
      if (tc::elect_one()) {
#pragma unroll 1
        for (int c = 0; c < kc; ++c) {
          const uint64_t adesc = tc::umma_desc_k_sw128(sa + c * A_BYTES);
          const uint64_t bdesc = tc::umma_desc_k_sw128(sa + KC * A_BYTES + c * b_bytes);
#pragma unroll
          for (int kk = 0; kk < BK / 16; ++kk)  // +32 bytes (=2 in descriptor units) per K=16 slice inside the swizzle atom
            tc::umma_ss(tmem_d, adesc + 2 * kk, bdesc + 2 * kk, idesc, (i | c | kk) != 0);
        }
        tc::umma_commit(bar_empty + 8 * ring);  // frees the smem slot when these MMAs retire
      }
      __syncwarp();
      ring = ring + 1 == stages ? 0 : ring + 1;
      k.next(g, kc);
    }
    if (tc::elect_one()) tc::umma_commit(bar_acc_full + 8 * ab);
    __syncwarp();
    ++acc_n;
  }
  st.full_par = full_par;
}

// epilogue warps 2..5: each thread owns one accumulator row (TMEM lane).  32 columns per rolled iteration; the TMEM load of
// chunk c+1 is issued as soon as chunk c has been moved out of the load registers (bias add), so the ~0.25 us TMEM
// round trip overlaps the conversion / stores of the previous chunk (serialised it made a 160-wide tile take 3 us).
__device__ __noinline__ void gemm_epilogue(const SeqGemm& g, uint8_t* smem, uint32_t smem_base, uint32_t tmem_base,
                                           RoleState& st, uint32_t acc_n) {
  long long* const tr = st.tr;
  const long long t0 = st.t0;
  // hot descriptor fields in registers too (the asm memory clobbers would re-read them from shared memory)
  const int gM = g.M, gN = g.N, ghw = g.hw, gldo = g.ldo, gsplits = g.splits;
  float* const gpartial = g.partial;
  int* const gcounters = g.counters;
  const float* const gbias = g.bias;
  const float* const growvec = g.rowvec;
  void* const gout = g.out;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int BN = g.bn;
  const uint32_t stage_bytes = KC * (A_BYTES + (uint32_t)BN * 128u);
  const uint32_t bar_acc_full = smem_base + OFF_ACC_FULL, bar_acc_empty = smem_base + OFF_ACC_EMPTY;
  const int num_work = g.mt * g.nt * gsplits;
  const int q = warp & 3;  // TMEM lane quarter this warp may access
  const int row = q * 32 + lane;
  // solo (at most one work item per CTA in this op): the producer and MMA warps are done by the time the accumulator is
  // ready, so all 8 warps run the epilogue - two per TMEM lane quarter, alternating 32-column chunks.  One warp per
  // scheduler issues an instruction every ~2.7 cycles (measured), which is what bounds this code.
  const bool solo = g.solo != 0;
  const int NE = solo ? SEQ_THREADS : 128;            // epilogue threads
  const int et = solo ? (int)threadIdx.x : (int)threadIdx.x - 64;
  const int nh = solo ? 2 : 1, h = solo ? warp >> 2 : 0;  // warps per quarter, this warp's index among them
  float* colvec = reinterpret_cast<float*>(smem + AUX_BYTES + g.stages * stage_bytes);  // [image in tile][BN]
  const int mode = g.mode;
  const int imgs_in_tile = (BM + ghw - 1) / ghw;
  const bool table = (gbias || growvec) && imgs_in_tile <= CV_IMGS;  // else (sub-4x4 maps): bias / rowvec straight from L2
  // One item per CTA: once the accumulator is ready the shared-memory ring is idle, so the bf16 tile is assembled there
  // (row per thread, 16-byte padded pitch: conflict-free) and leaves as whole 128-byte lines.  A row-per-thread global
  // store touches 32 different lines per instruction, which made the stores of a 128 x 160 tile take ~2 us.
  const bool staged = g.stage_out != 0 && gpartial == nullptr;
  const int out_cols = mode == 1 ? BN / 2 : BN;
  const int TP = out_cols + 8;
  bf16* tile = reinterpret_cast<bf16*>(smem + AUX_BYTES);
#pragma unroll 1
  for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
    const int mtile = w % g.mt, ntile = (w / g.mt) % g.nt, z = w / (g.mt * g.nt);
    const int m0 = mtile * BM, n0 = ntile * BN;
    const uint32_t ab = acc_n & 1u;
    const uint32_t trow = tmem_base + ab * ACC_COLS + ((uint32_t)(q * 32) << 16);
    const int m = m0 + row;
    const bool ok = m < gM;
    const int img = m / ghw, img0 = m0 / ghw;
    // ---- while the main loop of this item runs: warps 2..5 stage everything the epilogue needs that is not the accumulator
    const bool use_res = !gpartial && mode == 0 && g.residual != nullptr;
    const int e4 = (int)threadIdx.x - 64;  // 0..127 among the four staging warps
    if (warp >= 2 && warp < 6) {
      asm volatile("bar.sync 1, 128;" ::: "memory");  // previous item's readers of the table are done
      if (table) {  // bias[n] + rowvec[image, n] for the images of this tile
#pragma unroll 1
        for (int i = e4; i < imgs_in_tile * BN; i += 128) {
          const int b = i / BN, c = i - b * BN;
          float v = gbias ? gbias[n0 + c] : 0.f;
          if (growvec && (int64_t)(img0 + b) * ghw < gM) v += __ldcg(growvec + (int64_t)(img0 + b) * g.rowvec_ld + n0 + c);
          colvec[b * BN + c] = v;
        }
      }
      // residual row (bf16): all loads in flight now, parked as packed pairs in the spare TMEM columns [BN, BN + BN/2) of
      // this accumulator buffer (the MMAs write [0, BN)); the rolled chunk loop reads them back with a dynamic TMEM address
      if (use_res) {
        const uint4* rr = reinterpret_cast<const uint4*>(g.residual + (int64_t)min(m, gM - 1) * g.res_ld + n0);
        uint4 rv[20];
#pragma unroll
        for (int c = 0; c < 20; ++c)
          if (c * 8 < BN) rv[c] = __ldcg(rr + c);
#pragma unroll
        for (int c = 0; c < 10; ++c)
          if (c * 16 < BN) tmem_st8(trow + BN + c * 8, rv[2 * c], rv[2 * c + 1]);
        tc::tmem_st_wait();
        tc::tc_fence_before();  // solo: the other warp of this quarter reads half of these columns back
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
    }
    if (solo) asm volatile("bar.sync 2, %0;" ::"n"(SEQ_THREADS) : "memory");  // all hands: producer / MMA warps have finished their loops
    const float* cvrow = colvec + (table ? img - img0 : 0) * BN;
    if (tr && et == 0) tr[8] = clock64() - t0;  // staged, waiting for the accumulator
    tc::mbar_wait(bar_acc_full + 8 * ab, (acc_n >> 1) & 1u);
    tc::tc_fence_after();
    if (tr && et == 0) tr[9] = clock64() - t0;  // accumulator ready
    uint32_t r[32], rs[16];
    const int cstep = 32 * nh;
    if (h * 32 < BN) {
      tc::tmem_ld32(trow + h * 32, r);
      if (use_res) tc::tmem_ld16(trow + BN + h * 16, rs);
    }
#pragma unroll 1
    for (int c0 = h * 32; c0 < BN; c0 += cstep) {
      const int n = n0 + c0;
      const bool more = c0 + cstep < BN;
      tmem_ld_wait_dep(r);
      if (use_res) tmem_ld_dep16(rs);
      if (gpartial) {  // split-K: raw fp32 partial, reduced (+ epilogue) below or by the SEQ_SPLITK_REDUCE op
        if (ok) {
          float* pp = gpartial + ((int64_t)z * gM + m) * gN + n;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_global_v8(pp + 8 * j, r[8 * j], r[8 * j + 1], r[8 * j + 2], r[8 * j + 3], r[8 * j + 4], r[8 * j + 5], r[8 * j + 6],
                         r[8 * j + 7]);
        }
        if (more) tc::tmem_ld32(trow + c0 + cstep, r);
        if (tr && et == 0 && c0 == 0) tr[13] = clock64() - t0;  // first chunk done
        continue;
      }
      float v[32];
      if (table) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 b = *reinterpret_cast<const float4*>(cvrow + c0 + j);  // smem, same address across the warp's rows of one image
          v[j] = __uint_as_float(r[j]) + b.x; v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
          v[j + 2] = __uint_as_float(r[j + 2]) + b.z; v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
      }
      uint32_t rq[16];
      if (use_res) {
#pragma unroll
        for (int e = 0; e < 16; ++e) rq[e] = rs[e];
      }
      if (more) {  // the load registers are free again: next chunk in flight while this one is converted and stored
        tc::tmem_ld32(trow + c0 + cstep, r);
        if (use_res) tc::tmem_ld16(trow + BN + ((c0 + cstep) >> 1), rs);
      }
      if (tr && et == 0 && c0 == 0) tr[13] = clock64() - t0;  // first chunk out of the load registers
      if (!ok) continue;
      if (!table && (gbias || growvec)) {  // rare: more than CV_IMGS images per tile (maps smaller than 4x4)
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (gbias) {
            const float4 b = *reinterpret_cast<const float4*>(gbias + n + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
          if (growvec) {
            const float4 b = __ldcg(reinterpret_cast<const float4*>(growvec + (int64_t)img * g.rowvec_ld + n + j));
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
      }
      if (mode == 1) {
        // GEGLU: columns [16i, 16i+8) = values, [16i+8, 16i+16) = gates of the same 8 hidden channels -> 16 outputs
        float o[16];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) o[8 * i + j] = v[16 * i + j] * gelu_exact(v[16 * i + 8 + j]);
        if (staged) {
          uint4* tp = reinterpret_cast<uint4*>(tile + row * TP + (c0 >> 1));
          tp[0] = pack8(o);
          tp[1] = pack8(o + 8);
        } else {
          st_global_v8(reinterpret_cast<bf16*>(gout) + (int64_t)m * gldo + (n >> 1), pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]),
                       pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]), pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]),
                       pack_bf16(o[12], o[13]), pack_bf16(o[14], o[15]));
        }
      } else if (mode == 4) {
        float* op = reinterpret_cast<float*>(gout) + (int64_t)m * gldo + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_global_v8(op + 8 * j, __float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]), __float_as_uint(v[8 * j + 2]),
                       __float_as_uint(v[8 * j + 3]), __float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]),
                       __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]));
      } else if (mode == 2) {  // fp32 NCHW head
        const int pix = m - img * ghw;
        float* op = reinterpret_cast<float*>(gout) + (int64_t)img * g.n_valid * ghw + pix;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n + j < g.n_valid) op[(int64_t)(n + j) * ghw] = v[j];
      } else {  // 0: bf16 (+ residual); 3: SiLU then bf16
        if (use_res) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const float2 f = unpack_bf16(rq[e]);
            v[2 * e] += f.x;
            v[2 * e + 1] += f.y;
          }
        }
        if (mode == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
        }
        if (staged) {
          uint4* tp = reinterpret_cast<uint4*>(tile + row * TP + c0);
#pragma unroll
          for (int j = 0; j < 4; ++j) tp[j] = pack8(v + 8 * j);
        } else {
          bf16* op = reinterpret_cast<bf16*>(gout) + (int64_t)m * gldo + n;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            st_global_v8(op + 16 * j, pack_bf16(v[j * 16], v[j * 16 + 1]), pack_bf16(v[j * 16 + 2], v[j * 16 + 3]),
                         pack_bf16(v[j * 16 + 4], v[j * 16 + 5]), pack_bf16(v[j * 16 + 6], v[j * 16 + 7]),
                         pack_bf16(v[j * 16 + 8], v[j * 16 + 9]), pack_bf16(v[j * 16 + 10], v[j * 16 + 11]),
                         pack_bf16(v[j * 16 + 12], v[j * 16 + 13]), pack_bf16(v[j * 16 + 14], v[j * 16 + 15]));
        }
      }
    }
    if (staged) {  // the assembled tile -> global, consecutive threads on consecutive 16-byte pieces of a row
      asm volatile("bar.sync 3, %0;" ::"r"(NE) : "memory");
      if (tr && et == 0 && !gcounters) tr[15] = clock64() - t0;
      const int PV = out_cols / 8;
      const int rows_valid = min(BM, gM - m0);
      bf16* obase = reinterpret_cast<bf16*>(gout) + (int64_t)m0 * gldo + (mode == 1 ? n0 / 2 : n0);
#pragma unroll 2
      for (int i = et; i < rows_valid * PV; i += NE) {
        const int rr = i / PV, c8 = i - rr * PV;
        *reinterpret_cast<uint4*>(obase + (int64_t)rr * gldo + c8 * 8) = *reinterpret_cast<const uint4*>(tile + rr * TP + c8 * 8);
      }
    }
    tc::tc_fence_before();
    if (warp >= 2 && warp < 6) tc::mbar_arrive(bar_acc_empty + 8 * ab);  // this thread is done reading the accumulator buffer
    ++acc_n;
    if (tr && et == 0) tr[10] = clock64() - t0;  // epilogue stores issued
    if (gcounters) {
      // ---- split-K reduction fused into the op: all splits of a tile are co-resident (one work item per CTA), so they
      // meet at a global counter; each then reduces 1/splits of the tile's rows in fixed z order (bit-stable) and applies
      // the epilogue.  The slices of all splits are pulled into the (by now idle) shared-memory ring with cp.async, i.e.
      // every 16-byte piece is in flight at once: one L2 round trip instead of one per split.
      const int tile = mtile + ntile * g.mt;
      asm volatile("bar.sync 3, %0;" ::"r"(NE) : "memory");
      if (et == 0) {  // one cumulative fence publishes the 128 threads' partial rows; acquire on the way out
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        atomicAdd(&gcounters[tile], 1);
        uint32_t spins = 0;
        while (ld_acquire_u32(reinterpret_cast<const unsigned*>(&gcounters[tile])) < (unsigned)gsplits) {
          if (++spins > (1u << 25)) __trap();
        }
      }
      asm volatile("bar.sync 3, %0;" ::"r"(NE) : "memory");
      if (tr && et == 0) tr[11] = clock64() - t0;  // all splits arrived
      const int rows_per = (BM + gsplits - 1) / gsplits;
      const int r0 = z * rows_per, r1 = min(min(BM, r0 + rows_per), gM - m0);
      const int NV4 = BN / 4;                       // 16-byte pieces per row
      const int per_split = max(r1 - r0, 0) * NV4;  // pieces of one split's slice
      float4* stage = reinterpret_cast<float4*>(smem + AUX_BYTES);  // [split][row][BN] fp32
      const int64_t zstride = (int64_t)gM * gN;
      const float* pbase = gpartial + (int64_t)(m0 + r0) * gN + n0;
#pragma unroll 1
      for (int i = et; i < per_split; i += NE) {  // this thread's 16-byte piece of the slice, from every split
        const int rr = i / NV4, c4 = i - rr * NV4;
        const float* src = pbase + (int64_t)rr * gN + c4 * 4;
        float4* dst = stage + i;
#pragma unroll 4
        for (int zz = 0; zz < gsplits; ++zz) cp_async16(dst + zz * per_split, src + zz * zstride);
      }
      if (tr && et == 0) tr[14] = clock64() - t0;  // slice copies issued
      cp_async_wait_all();
      asm volatile("bar.sync 3, %0;" ::"r"(NE) : "memory");
      if (tr && et == 0) tr[15] = clock64() - t0;  // slices in shared memory
      const int NV = BN / 8;
#pragma unroll 1
      for (int i = et; i < max(r1 - r0, 0) * NV; i += NE) {
        const int rr = i / NV, nn = (i - rr * NV) * 8;
        const int mm = m0 + r0 + rr;
        const uint4 res = g.residual ? ld_cg16(g.residual + (int64_t)mm * g.res_ld + n0 + nn) : make_uint4(0u, 0u, 0u, 0u);
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = 0.f;
        const float4* sp = stage + rr * NV4 + (nn >> 2);
#pragma unroll 2
        for (int zz = 0; zz < gsplits; ++zz) {  // fixed z order: bit-stable
          const float4 a = sp[zz * per_split], b = sp[zz * per_split + 1];
          v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
          v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        if (table) {
          const float* cr = colvec + (mm / ghw - img0) * BN + nn;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += cr[j];
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            if (gbias) v[j] += gbias[n0 + nn + j];
            if (growvec) v[j] += __ldcg(growvec + (int64_t)(mm / ghw) * g.rowvec_ld + n0 + nn + j);
          }
        }
        float f[8];
        unpack8(res, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += f[j];
        *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(gout) + (int64_t)mm * gldo + n0 + nn) = pack8(v);
      }
      if (tr && et == 0) tr[12] = clock64() - t0;  // slice reduced
      asm volatile("bar.sync 3, %0;" ::"r"(NE) : "memory");
      if (et == 0) {  // the last split to finish re-arms the counters for the next op
        if (atomicAdd(&gcounters[g.mt * g.nt + tile], 1) == gsplits - 1) {
          gcounters[tile] = 0;
          gcounters[g.mt * g.nt + tile] = 0;
        }
      }
    }
  }
}

// =====================================================================================================
// SEQ_GN  GroupNorm (+SiLU), two-pass (mean, then centred second moment) so that |mean| >> std costs no precision
// =====================================================================================================
// One item = (image, channel block of whole groups, pixel chunk), owned by `gw` warps (1, 2, 4 or 8: small images pack
// several items into a CTA).  The item is copied to shared memory with cp.async (all copies in flight, no registers),
// then every pass is a short rolled loop over shared memory: thread t of the item's group owns vector column cv (8
// channels, at most two groups) and pixels pl, pl + lanes, ...  With ps > 1 (big images split over CTAs) the chunk
// statistics (mean, M2) meet in global memory across one grid barrier and are merged with Chan's formula in chunk order,
// while the chunk stays in shared memory for the normalisation.  Fixed reduction orders: bit-stable.
__device__ __noinline__ unsigned gn_op(const SeqGN& p, uint8_t* smem, GridBarrier gb) {
  float* red = reinterpret_cast<float*>(smem + OFF_RED);    // [warp][4]
  float* stat = reinterpret_cast<float*>(smem + OFF_STAT);  // [group of the CTA][4][2]: mean, rstd
  const int C = p.c0 + p.c1, NV = p.cb / 8, GB = p.cb / p.cgn, nblk = C / p.cb;
  const int TG = p.gw * 32, groups = NWARPS / p.gw;          // threads per item, items per CTA and round
  const int gid = threadIdx.x / TG, tg = threadIdx.x - gid * TG;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, w0 = gid * p.gw;
  const int lanes = TG / NV, cv = tg % NV, pl = tg / NV;     // pixel lanes of the item; this thread's column and first pixel
  const bool active = pl < lanes;
  const int g_lo = (cv * 8) / p.cgn, eb = (g_lo + 1) * p.cgn - cv * 8;  // channels e < eb of the vector -> g_lo, the rest g_lo + 1
  const int g_hi = min(g_lo + 1, GB - 1);
  const int slab_bytes = p.px * p.cb * 2;
  const float inv_cnt = 1.f / ((float)p.px * (float)p.cgn);
  const int per_round = gridDim.x * groups;
  const int rounds = (p.n_items + per_round - 1) / per_round;

  // sum `v` (this thread's contribution to groups g_lo / g_lo+1) over the item's threads: v_lo, v_hi -> totals of g_lo, g_hi
  auto group_sum = [&](float& v_lo, float& v_hi) {
    float S[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) S[g] = warp_sum((g == g_lo ? v_lo : 0.f) + (g == g_lo + 1 ? v_hi : 0.f));
    __syncthreads();  // previous readers of `red` are done
    if (lane == 0) *reinterpret_cast<float4*>(red + warp * 4) = make_float4(S[0], S[1], S[2], S[3]);
    __syncthreads();
    float T[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
    for (int w = 0; w < p.gw; ++w) {  // the item's warps, in order
      const float4 r = *reinterpret_cast<const float4*>(red + (w0 + w) * 4);
      T[0] += r.x; T[1] += r.y; T[2] += r.z; T[3] += r.w;
    }
    v_lo = v_hi = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      v_lo = g == g_lo ? T[g] : v_lo;
      v_hi = g == g_hi ? T[g] : v_hi;
    }
  };

#pragma unroll 1
  for (int phase = 0; phase < (p.ps > 1 ? 2 : 1); ++phase) {
    if (phase == 1) gb.n = grid_sync(gb.ctr, gb.n, gb.flags, gb.timing, gb.trace, gb.op, nullptr, nullptr);
#pragma unroll 1
    for (int k = 0; k < rounds; ++k) {
      const int item = k * per_round + gid * gridDim.x + blockIdx.x;
      const bool valid = item < p.n_items && active;
      const int it = min(item, p.n_items - 1);
      const int chunk = it % p.ps, blk = (it / p.ps) % nblk, img = it / (p.ps * nblk);
      const int ch = blk * p.cb + cv * 8;
      const int64_t pix0 = (int64_t)img * p.hw + (int64_t)chunk * p.px;
      const bool first = ch < p.c0;
      const bf16* src = first ? p.x0 + pix0 * p.c0 + ch : p.x1 + pix0 * p.c1 + (ch - p.c0);
      const int pitch = first ? p.c0 : p.c1;
      const uint4* slab = reinterpret_cast<const uint4*>(smem + AUX_BYTES + (size_t)(k * groups + gid) * slab_bytes) + cv;
      float m_lo, m_hi, r_lo, r_hi;
      if (phase == 0) {
        if (valid) {
#pragma unroll 1
          for (int pp = pl; pp < p.px; pp += lanes) cp_async16(const_cast<uint4*>(slab) + pp * NV, src + (int64_t)pp * pitch);
        }
        cp_async_wait_all();
        __syncthreads();
        // ---- pass 1: per-channel sums -> group means
        float sa[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) sa[e] = 0.f;
        if (valid) {
#pragma unroll 2
          for (int pp = pl; pp < p.px; pp += lanes) {
            float f[8];
            unpack8(slab[pp * NV], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) sa[e] += f[e];
          }
        }
        m_lo = m_hi = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          m_lo += e < eb ? sa[e] : 0.f;
          m_hi += e < eb ? 0.f : sa[e];
        }
        group_sum(m_lo, m_hi);
        m_lo *= inv_cnt;
        m_hi *= inv_cnt;
        // ---- pass 2: centred second moment
        r_lo = r_hi = 0.f;
        if (valid) {
#pragma unroll 2
          for (int pp = pl; pp < p.px; pp += lanes) {
            float f[8];
            unpack8(slab[pp * NV], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const bool lo = e < eb;
              const float d = f[e] - (lo ? m_lo : m_hi);
              r_lo = lo ? fmaf(d, d, r_lo) : r_lo;
              r_hi = lo ? r_hi : fmaf(d, d, r_hi);
            }
          }
        }
        group_sum(r_lo, r_hi);
        if (p.ps > 1) {  // (mean, M2) of this chunk per group; merged after the grid barrier
          if (valid && pl == 0) {  // one thread per vector column reports the group(s) its channels belong to (every group is
            float* po = p.partial + (int64_t)item * 8;  // some column's first or second group; duplicates store equal values)
            *reinterpret_cast<float2*>(po + 2 * g_lo) = make_float2(m_lo, r_lo);
            if (eb < 8 && g_lo + 1 < GB) *reinterpret_cast<float2*>(po + 2 * (g_lo + 1)) = make_float2(m_hi, r_hi);
          }
          continue;
        }
        r_lo = rsqrtf(r_lo * inv_cnt + p.eps);
        r_hi = rsqrtf(r_hi * inv_cnt + p.eps);
      } else {
        // ---- merge the image's chunks in chunk order (equal counts): mean of means, M2 += n (mean_c - mean)^2
        __syncthreads();  // previous round's readers of `stat` are done
        if (item < p.n_items && tg < GB) {
          const float* pi = p.partial + (int64_t)(item - chunk) * 8 + 2 * tg;
          float msum = 0.f;
#pragma unroll 1
          for (int c = 0; c < p.ps; ++c) msum += __ldcg(pi + c * 8);
          const float mean = msum / (float)p.ps;
          const float n_c = (float)p.px * (float)p.cgn;
          float m2 = 0.f;
#pragma unroll 1
          for (int c = 0; c < p.ps; ++c) {
            const float2 pr = __ldcg(reinterpret_cast<const float2*>(pi + c * 8));
            const float d = pr.x - mean;
            m2 += pr.y + n_c * d * d;
          }
          stat[(gid * 4 + tg) * 2] = mean;
          stat[(gid * 4 + tg) * 2 + 1] = rsqrtf(m2 / (n_c * (float)p.ps) + p.eps);
        }
        __syncthreads();
        m_lo = stat[(gid * 4 + g_lo) * 2];
        r_lo = stat[(gid * 4 + g_lo) * 2 + 1];
        m_hi = stat[(gid * 4 + g_hi) * 2];
        r_hi = stat[(gid * 4 + g_hi) * 2 + 1];
      }
      // ---- normalise from shared memory
      if (valid) {
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
        bf16* dst = p.out + pix0 * C + ch;
#pragma unroll 2
        for (int pp = pl; pp < p.px; pp += lanes) {
          float f[8];
          unpack8(slab[pp * NV], f);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float y = fmaf(f[e], sc[e], sh[e]);
            f[e] = p.silu ? silu_f(y) : y;
          }
          *reinterpret_cast<uint4*>(dst + (int64_t)pp * C) = pack8(f);
        }
      }
    }
  }
  return gb.n;
}

// =====================================================================================================
// SEQ_LN  LayerNorm over the channel dim: one warp per token, R tokens staged in shared memory per warp and pass
// =====================================================================================================
__device__ __noinline__ void ln_op(const SeqLN& p, uint8_t* smem) {
  constexpr int R = 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * NWARPS + warp, GW = gridDim.x * NWARPS;
  const int nv = p.c / 8;
  const float inv_c = 1.f / (float)p.c;
  // gamma / beta once per CTA, rows per warp behind them
  float* gam = reinterpret_cast<float*>(smem + AUX_BYTES);
  float* bet = gam + p.c;
  uint4* rows = reinterpret_cast<uint4*>(bet + p.c) + (size_t)warp * R * nv;
#pragma unroll 1
  for (int i = threadIdx.x; i < p.c / 4; i += SEQ_THREADS) {
    cp_async16(gam + 4 * i, p.gamma + 4 * i);
    cp_async16(bet + 4 * i, p.beta + 4 * i);
  }
  auto stage_rows = [&](int r0) {
#pragma unroll 1
    for (int u = 0; u < R; ++u) {
      const int row = r0 + u * GW;
      if (row < p.rows) {
#pragma unroll 1
        for (int v = lane; v < nv; v += 32) cp_async16(rows + u * nv + v, p.x + (int64_t)row * p.c + v * 8);
      }
    }
  };
  stage_rows(gw);  // first pass of this warp (if any), in flight together with gamma / beta
  cp_async_wait_all();
  __syncthreads();
#pragma unroll 1
  for (int r0 = gw; r0 < p.rows; r0 += R * GW) {
    if (r0 != gw) {
      stage_rows(r0);
      cp_async_wait_all();
      __syncwarp();
    }
#pragma unroll 1
    for (int u = 0; u < R; ++u) {
      const int row = r0 + u * GW;
      if (row >= p.rows) break;  // warp-uniform
      const uint4* xr = rows + u * nv;
      float s = 0.f;
#pragma unroll 1
      for (int v = lane; v < nv; v += 32) {
        float f[8];
        unpack8(xr[v], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s += f[j];
      }
      const float mean = warp_sum(s) * inv_c;
      float q = 0.f;
#pragma unroll 1
      for (int v = lane; v < nv; v += 32) {
        float f[8];
        unpack8(xr[v], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = f[j] - mean;
          q = fmaf(d, d, q);
        }
      }
      const float rstd = rsqrtf(warp_sum(q) * inv_c + p.eps);
#pragma unroll 1
      for (int v = lane; v < nv; v += 32) {
        float f[8];
        unpack8(xr[v], f);
        const float4 g0 = *reinterpret_cast<const float4*>(gam + v * 8), g1 = *reinterpret_cast<const float4*>(gam + v * 8 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bet + v * 8), b1 = *reinterpret_cast<const float4*>(bet + v * 8 + 4);
        f[0] = (f[0] - mean) * rstd * g0.x + b0.x; f[1] = (f[1] - mean) * rstd * g0.y + b0.y;
        f[2] = (f[2] - mean) * rstd * g0.z + b0.z; f[3] = (f[3] - mean) * rstd * g0.w + b0.w;
        f[4] = (f[4] - mean) * rstd * g1.x + b1.x; f[5] = (f[5] - mean) * rstd * g1.y + b1.y;
        f[6] = (f[6] - mean) * rstd * g1.z + b1.z; f[7] = (f[7] - mean) * rstd * g1.w + b1.w;
        *reinterpret_cast<uint4*>(p.out + (int64_t)row * p.c + v * 8) = pack8(f);
      }
    }
    __syncwarp();  // the row buffers are refilled by the next pass
  }
}

// =====================================================================================================
// small elementwise ops
// =====================================================================================================
__device__ __noinline__ void upsample_op(const SeqEW& p) {  // nearest 2x, bf16 NHWC; 32-bit index arithmetic
  const bf16* x = reinterpret_cast<const bf16*>(p.src);
  bf16* out = reinterpret_cast<bf16*>(p.dst);
  const int cv = p.c / 8, w2 = 2 * p.w, h2 = 2 * p.h;
  const int total = p.n_img * h2 * w2 * cv;
#pragma unroll 1
  for (int i = blockIdx.x * SEQ_THREADS + threadIdx.x; i < total; i += gridDim.x * SEQ_THREADS) {
    const int v = i % cv, q = i / cv;
    const int ox = q % w2, q2 = q / w2;
    const int oy = q2 % h2, img = q2 / h2;
    *reinterpret_cast<uint4*>(out + (int64_t)i * 8) = ld_cg16(x + ((int64_t)(img * p.h + oy / 2) * p.w + ox / 2) * p.c + v * 8);
  }
}

// conv_in operand: explicit im2col of the fp32 NCHW denoiser input.
// out[m, tap*cin + c] = latents[img, c, y+r-1, x+s-1] (zero outside / for k >= 9*cin), m = (img, y, x)
__device__ __noinline__ void im2col_op(const SeqEW& p) {
  const float* x = reinterpret_cast<const float*>(p.src);
  bf16* out = reinterpret_cast<bf16*>(p.dst);
  const int cin = p.c, h = p.h, w = p.w, kpad = p.aux;
  const int total = p.n_img * h * w * kpad;  // one thread = one element: consecutive threads write consecutive bf16
#pragma unroll 1
  for (int i = blockIdx.x * SEQ_THREADS + threadIdx.x; i < total; i += gridDim.x * SEQ_THREADS) {
    const int m = i / kpad, k = i - m * kpad;
    const int px = m % w, py = (m / w) % h, img = m / (w * h);
    float v = 0.f;
    if (k < 9 * cin) {
      const int tap = k / cin, c = k - tap * cin;
      const int yy = py + tap / 3 - 1, xx = px + tap % 3 - 1;
      if (yy >= 0 && yy < h && xx >= 0 && xx < w) v = __ldcg(x + ((img * cin + c) * h + yy) * w + xx);
    }
    out[i] = __float2bfloat16(v);
  }
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin] -> bf16 (the operand autocast feeds
// time_embedding.linear_1 in the reference)
__device__ __noinline__ void sinusoid_op(const SeqEW& p) {
  const int64_t* t = reinterpret_cast<const int64_t*>(p.src);
  bf16* out = reinterpret_cast<bf16*>(p.dst);
  const int n = p.n_img, dim = p.c, half = dim / 2;
#pragma unroll 1
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
  const int total = M * nv;
#pragma unroll 1
  for (int i = blockIdx.x * SEQ_THREADS + threadIdx.x; i < total; i += gridDim.x * SEQ_THREADS) {
    const int m = i / nv, n = (i - m * nv) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int z = 0; z < splits; ++z) {
      const float4* pp = reinterpret_cast<const float4*>(partial + ((int64_t)z * M + m) * N + n);
      const float4 a = __ldcg(pp), b = __ldcg(pp + 1);
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (p.bias) v[j] += p.bias[n + j];
      if (p.rowvec) v[j] += __ldcg(p.rowvec + (int64_t)(m / hw) * rowvec_ld + n + j);
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
                                                             long long* timing, int flags, long long* trace) {
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

  // The tensor maps live in global memory (written by cudaMemcpy before the launch) and are read through the tensormap
  // proxy: every CTA acquires them before first use (CUDA programming guide, "tensor map in global memory") - all maps of
  // the launch here, spread over the threads, so that no op pays for it on its critical path.
  if (flags & SEQ_F_TMAP_FENCE) {
    constexpr int MAPS = (MVLDM_MAX_SEGS + 1) * KC;
#pragma unroll 1
    for (int i = threadIdx.x; i < n_ops * MAPS; i += SEQ_THREADS) {
      const CUtensorMap* m = &ops[i / MAPS].tmA[0][0] + (i % MAPS);
      asm volatile("fence.proxy.tensormap::generic.acquire.sys [%0], 128;" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
    }
    __syncthreads();
  }
  // the op descriptors and tensor maps are touched once per forward (cold in L2 behind 1.5 GB of weights): pull the whole
  // list into L2 now, the first op's maps into the descriptor cache
  {
    const char* base = reinterpret_cast<const char*>(ops);
    const int lines = n_ops * (int)(sizeof(SeqOp) / 128);
#pragma unroll 1
    for (int i = threadIdx.x + (blockIdx.x % 4) * SEQ_THREADS; i < lines; i += 4 * SEQ_THREADS)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)i * 128));
    if (threadIdx.x == 0 && op.type == SEQ_GEMM) prefetch_tensormaps(ops, op.g.nseg);
  }
  GridBarrier gb{sync, 0u, timing, flags, trace, 0};
  RoleState st{0xffffffffu, 0u, nullptr, 0};
  uint32_t acc_base = 0;  // accumulator hand-offs of this CTA so far (buffer = n & 1, phase = (n >> 1) & 1): every thread counts
#pragma unroll 1
  for (int i = 0; i < n_ops; ++i) {
    const int type = op.type;
    gb.op = i;
    st.tr = trace ? trace + ((size_t)gb.n * gridDim.x + blockIdx.x) * 16 : nullptr;
    st.t0 = clock64();
    if (type == SEQ_GEMM) {
      const bool solo = op.g.solo != 0;
      if (warp == 0) {
        if (lane == 0) {
          gemm_producer(op.g, ops + i, smem_base, st);
          // the rest of this op (MMA tail, epilogue, reduction, barrier) hides the HBM latency of the next op's weights
          if (op.next_nseg > 0) prefetch_tensormaps(ops + i + 1, op.next_nseg);
          if (flags & SEQ_F_PREFETCH) prefetch_next_weights(op.pf);
        }
        __syncwarp();
      } else if (warp == 1) {
        gemm_mma(op.g, smem_base, tmem_base, st, acc_base);
      }
      if (solo || (warp >= 2 && warp < 6)) gemm_epilogue(op.g, smem, smem_base, tmem_base, st, acc_base);
      {  // every thread counts this CTA's accumulator hand-offs of the op
        const int nw = op.g.mt * op.g.nt * op.g.splits;
        acc_base += (int)blockIdx.x < nw ? (uint32_t)((nw - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) : 0u;
      }
    } else {
      if (threadIdx.x == 0 && op.next_nseg > 0) prefetch_tensormaps(ops + i + 1, op.next_nseg);
      if (type == SEQ_GN) gb.n = gn_op(op.gn, smem, gb);
      else if (type == SEQ_LN) ln_op(op.ln, smem);
      else if (type == SEQ_UPSAMPLE) upsample_op(op.ew);
      else if (type == SEQ_IM2COL) im2col_op(op.ew);
      else if (type == SEQ_SINUSOID) sinusoid_op(op.ew);
      else if (type == SEQ_SPLITK_REDUCE) splitk_reduce_op(op.ew);
    }
    if (i + 1 < n_ops) gb.n = grid_sync(gb.ctr, gb.n, gb.flags, gb.timing, gb.trace, i, ops + i + 1, smem + OFF_DESC);
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

// ONE GEMM op per launch: same roles, the op descriptor and its tensor maps in parameter (constant) space.
__global__ void __launch_bounds__(SEQ_THREADS, 1) seq_gemm_kernel(const __grid_constant__ SeqOp op, int flags) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = tc::smem_u32(smem_raw);
  const uint32_t smem_base = (raw_addr + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (smem_base - raw_addr);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const SeqGemm& g = op.c.g;
  if (threadIdx.x == 0) {
    prefetch_tensormaps(&op, g.nseg);
    for (int s = 0; s < MAX_STAGES; ++s) {
      tc::mbar_init(smem_base + OFF_BAR_FULL + 8 * s, 1);
      tc::mbar_init(smem_base + OFF_BAR_EMPTY + 8 * s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(smem_base + OFF_ACC_FULL + 8 * b, 1);
      tc::mbar_init(smem_base + OFF_ACC_EMPTY + 8 * b, 128);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<512>(smem_base + OFF_TMEM);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<const uint32_t*>(smem + OFF_TMEM);
  RoleState st{0xffffffffu, 0u, nullptr, 0};
  const bool solo = g.solo != 0;
  if (warp == 0) {
    if (lane == 0) {
      gemm_producer(g, &op, smem_base, st);
      if ((flags & SEQ_F_PREFETCH) && op.c.pf.bn > 0) {  // the next launch's weights -> L2 while this one drains
        SeqPrefetch pf = op.c.pf;
        pf.map = &op.tmPf;
        prefetch_next_weights(pf);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    gemm_mma(g, smem_base, tmem_base, st, 0u);
  }
  if (solo || (warp >= 2 && warp < 6)) gemm_epilogue(g, smem, smem_base, tmem_base, st, 0u);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<512>(tmem_base);
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
    if (d.residual && bn > 160) continue;  // the epilogue parks the residual row in the accumulator buffer's spare TMEM columns
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
  p.solo = work <= seq_grid() ? 1 : 0;
  p.stage_out = (p.solo && splits == 1 && (d.mode == 0 || d.mode == 1 || d.mode == 3)) ? 1 : 0;
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
  int next_gemm = -1;
  for (int i = n - 1; i >= 0; --i) {
    ops[i].c.index = i;
    ops[i].c.pf.map = nullptr;
    ops[i].c.pf.bn = 0;
    if (ops[i].c.type != SEQ_GEMM) continue;
    if (next_gemm >= 0) {
      const SeqGemm& g = ops[next_gemm].c.g;
      SeqPrefetch& pf = ops[i].c.pf;
      pf.map = dev_ops ? &dev_ops[next_gemm].tmB[KC - 1] : nullptr;
      ops[i].tmPf = ops[next_gemm].tmB[KC - 1];
      pf.bn = g.bn;
      pf.nt = g.nt;
      pf.nchunks = g.seg[g.nseg - 1].kchunk0 + g.seg[g.nseg - 1].ntaps * g.seg[g.nseg - 1].ncblk;
    }
    next_gemm = i;
  }
}

void seq_link_launch(SeqOp* ops, int n) {
  for (int i = 0; i < n; ++i) ops[i].c.next_nseg = (i + 1 < n && ops[i + 1].c.type == SEQ_GEMM) ? ops[i + 1].c.g.nseg : 0;
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
  const int64_t items0 = (int64_t)n_img * nblk;
  // big images on few (image, block) items: split the pixels over CTAs (statistics merged across a grid barrier)
  int ps = 1;
  while (items0 * ps * 2 <= grid && hw % (ps * 2) == 0 && hw / (ps * 2) >= 64) ps *= 2;
  p.ps = ps;
  p.px = hw / ps;
  p.n_items = (int)(items0 * ps);
  MV_CHECK((size_t)p.n_items * 8 <= seq_groupnorm_scratch_floats(n_img, C, groups) || ps == 1, "groupnorm: scratch too small");
  // warps per item: fewest rounds first (a round costs ~1 us of latency), then fewest vectors per thread
  const size_t slab = (size_t)p.px * cb * 2;
  int best_gw = 0;
  double best_cost = 1e30;
  for (int gw : {8, 4, 2, 1}) {
    if (ps > 1 && gw != 8) continue;
    if (gw * 32 < NV) continue;
    const int groups_cta = NWARPS / gw;
    const int64_t rounds = (p.n_items + (int64_t)grid * groups_cta - 1) / ((int64_t)grid * groups_cta);
    if ((size_t)rounds * groups_cta * slab > (size_t)WORK_BYTES) continue;
    const double vec_per_thread = (double)p.px / (double)(gw * 32 / NV);
    const double cost = rounds * (1.0 + 0.05 * vec_per_thread);
    if (cost < best_cost) {
      best_cost = cost;
      best_gw = gw;
    }
  }
  MV_CHECK(best_gw != 0, "groupnorm: image too large for the shared-memory slab (pixels per image x channel block)");
  p.gw = best_gw;
}

void seq_plan_layernorm(const bf16* x, int rows, int c, float eps, const float* gamma, const float* beta, bf16* out, SeqOp& op) {
  MV_CHECK(c % 8 == 0 && c <= 32 * 8 * 8 && (size_t)c * (8 + NWARPS * 4 * 2) <= (size_t)WORK_BYTES, "layernorm: unsupported channel count");
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

long long* g_seq_trace = nullptr;  // debug: device buffer the next launches write their barrier timeline to (or NULL)

void seq_configure() {  // once per device, outside any stream capture
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    MV_CUDA(cudaFuncSetAttribute(seq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    MV_CUDA(cudaFuncSetAttribute(seq_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
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
  static const int env_flags = [] {
    const char* e = getenv("MVLDM_SEQ_FLAGS");
    return e ? atoi(e) : 7;
  }();
  int flags = env_flags;
  long long* trace = g_seq_trace;
  if (g_seq_trace) g_seq_trace += (size_t)2 * n_ops * seq_grid() * 16;  // room for this launch's barriers (<= 2 per op)
  void* args[] = {(void*)&dev_ops, (void*)&n_ops, (void*)&sync, (void*)&timing, (void*)&flags, (void*)&trace};
  if (coop) {
    MV_CUDA(cudaLaunchCooperativeKernel((const void*)seq_kernel, dim3(seq_grid()), dim3(SEQ_THREADS), args, SMEM_BYTES, s));
  } else {
    MV_CUDA(cudaLaunchKernel((const void*)seq_kernel, dim3(seq_grid()), dim3(SEQ_THREADS), args, SMEM_BYTES, s));
  }
  MV_LAUNCHED();
}

void seq_launch_gemm(cudaStream_t s, const SeqOp& op) {
  MV_CHECK(op.c.type == SEQ_GEMM, "seq_launch_gemm: not a GEMM op");
  seq_configure();
  static const int env_flags = [] {
    const char* e = getenv("MVLDM_SEQ_FLAGS");
    return e ? atoi(e) : 7;
  }();
  const SeqGemm& g = op.c.g;
  const int work = g.mt * g.nt * g.splits;
  seq_gemm_kernel<<<dim3(std::min(work, seq_grid())), dim3(SEQ_THREADS), SMEM_BYTES, s>>>(op, env_flags);
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
  seq_link_launch(tmp.data(), n_ops);
  MV_CUDA(cudaMemcpyAsync(sl.ops, tmp.data(), sizeof(SeqOp) * n_ops, cudaMemcpyHostToDevice, s));
  seq_launch(s, sl.ops, n_ops, sl.sync, nullptr);
}

// debug / measurement: a launch of `n_ops` empty ops = the bare cost of an op boundary (tools/seq_barrier_bench.py)
void seq_debug_empty_ops(cudaStream_t s, int n_ops) {
  std::vector<SeqOp> ops(n_ops);
  for (auto& o : ops) seq_plan_layernorm(nullptr, 0, 8, 1e-5f, nullptr, nullptr, nullptr, o);
  seq_run_host_ops(s, ops.data(), n_ops);
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
