// tcgen05 / TMEM / TMA implicit-GEMM for sm_100a: conv3x3 (stride 1/2), 1x1 conv and Linear over bf16 NHWC
// activations, fp32 accumulation in tensor memory, fused epilogues (bias, per-image time-embedding row,
// residual add, GEGLU, fp32-NCHW head output).
//
//   D[128 x BN] (TMEM, fp32) += A[128 x 64] (smem, K-major SW128) * W[BN x 64]^T (smem, K-major SW128)
//
// A is never materialised: for every (segment, filter tap, 64-channel block) one TMA box
// [64 ch, bw, bh, bn] (bw*bh*bn = 128 output pixels) is fetched from the NHWC source at the tap's pixel
// offset; out-of-image coordinates are zero-filled by the TMA unit, which is exactly the conv's zero
// padding.  Stride-2 convs use the tensor map's traversal stride.  Up to three K-segments let one
// accumulator take conv2(3x3) + the 1x1 shortcut over the (possibly concatenated) block input.
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = TMEM owner + MMA issuer
// (one elected lane), warps 2..5 = epilogue (each owns the 32 TMEM lanes of its warp%4 quarter).
// Pipelines: smem full/empty ring of mbarriers (TMA <-> MMA), one accumulator-ready mbarrier (MMA -> epilogue).
// One CTA per SM (the smem ring takes ~220 KB); the accumulator is double-buffered in TMEM, so the epilogue of work item i
// overlaps the main loop of item i+1.
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
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

// SM count of the current device (one persistent CTA per SM)
int sm_count() {
  static int sms[kMaxDevices] = {};
  int dev = 0;
  MV_CUDA(cudaGetDevice(&dev));
  MV_CHECK(dev >= 0 && dev < kMaxDevices, "device index out of range");
  if (sms[dev] == 0) MV_CUDA(cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev));
  return sms[dev];
}

namespace {

constexpr int BM = 128, BK = 64;

constexpr int A_BYTES = BM * BK * 2;  // 16 KB

constexpr int KC = 2;  // 64-channel chunks per pipeline step: one TMA box per operand carries up to KC chunks

struct TcSeg {
  int ncblk, ntaps, stride;
  int spt;      // steps per tap = ceil(ncblk / KC)
  int kchunk0;  // index of this segment's first 64-wide K chunk in the weight matrix
  int dh[9], dw[9], coff[9];
};

struct TcParams {
  CUtensorMap tmA[MVLDM_MAX_SEGS][KC];  // [segment][chunks per box - 1]: 5-D (64 ch, w, h, image, chunk)
  CUtensorMap tmB[KC];                   // 3-D (64 k, n, chunk)
  TcSeg seg[MVLDM_MAX_SEGS];
  int nseg;
  int M, N, num_steps;
  int mt, nt, splits;   // work items = mt * nt * splits, m fastest
  int steps_per_split;  // pipeline steps per split (== num_steps when not split)
  float* partial;    // split-K: fp32 partial tiles [splits][M][N]; NULL when not split
  int* counters;     // split-K fused reduction: [2][mt*nt] arrive / done counters (all zero between launches), or NULL
  int hw, ow;        // output pixels per image / row width (tile -> image coordinates)
  // M tile = TMA box of bw pixels x bh rows x bn images (a_rows = bw*bh*bn <= 128 rows of the 128-row MMA; rows beyond
  // a_rows / beyond the image are never stored).  dense: the tile is 128 consecutive tokens (row r <-> token m0 + r), the
  // common case (widths dividing 128 with hw | 128 or 128 | hw); otherwise tiles walk (x block, y block, image block).
  int dense, bw, bh, bn, tx, ty, a_rows, oh, n_img;
  const float* bias;
  const float* rowvec;
  int rowvec_ld;
  const bf16* residual;
  int res_ld;
  int mode;
  void* out;
  int ldo, n_valid;
  const bf16* w;  // raw weight matrix [N, K] (L2 prefetch ahead of the dependency wait)
  int K;
  int opt;  // bit 0: bias/rowvec table in smem, bit 1: residual row prefetch, bit 2: 4-way unrolled fused reduce,
            // bit 3: L2-prefetch this CTA's first weight tile before griddepcontrol.wait, bit 4: release dependents at entry
};

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

// Persistent: grid = min(#work items, #SMs); every CTA walks work items w = blockIdx.x, +gridDim.x, ... where a work
// item is (m-tile, n-tile, k-split), m fastest so that the CTAs running concurrently share the same weight tile
// in L2.  The smem ring runs continuously across work items and the accumulator is double-buffered in TMEM, so
// the epilogue of item i (TMEM -> registers -> global) overlaps the main loop of item i+1.
// 256-bit global store: one instruction covers a full 32-byte sector per thread (rows are >= 64 B apart, so 16-byte
// stores would touch every sector twice)
__device__ __forceinline__ void st_global_v8(void* ptr, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4,
                                             uint32_t a5, uint32_t a6, uint32_t a7) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(a0), "r"(a1), "r"(a2), "r"(a3),
               "r"(a4), "r"(a5), "r"(a6), "r"(a7)
               : "memory");
}

// origin (first image, first row, first column) of M tile `mtile`
template <bool DENSE>
__device__ __forceinline__ void tile_origin(const TcParams& p, int mtile, int& img0, int& y0, int& x0) {
  if (DENSE) {
    const int m0 = mtile * BM;
    img0 = m0 / p.hw;
    y0 = (m0 - img0 * p.hw) / p.ow;
    x0 = 0;
  } else {
    const int ix = mtile % p.tx, iy = (mtile / p.tx) % p.ty;
    x0 = ix * p.bw;
    y0 = iy * p.bh;
    img0 = (mtile / (p.tx * p.ty)) * p.bn;
  }
}
// token (row of the [M, N] output) that accumulator row r of the tile holds, or -1; di = image index inside the tile
template <bool DENSE>
__device__ __forceinline__ int row_token(const TcParams& p, int mtile, int img0, int y0, int x0, int r, int& di) {
  if (DENSE) {
    const int m = mtile * BM + r;
    di = m / p.hw - img0;
    return m < p.M ? m : -1;
  }
  const int rx = r % p.bw, t = r / p.bw, ry = t % p.bh;
  di = t / p.bh;
  const int img = img0 + di, y = y0 + ry;
  return (r < p.a_rows && img < p.n_img && y < p.oh) ? (img * p.oh + y) * p.ow + x0 + rx : -1;
}

// DENSE: tiles of 128 consecutive tokens (every shape of the 32x32 / 16x16 / 64x64 configurations) - the index arithmetic of the
// general (columns x rows x images) box compiles away
template <int BN, int STAGES, bool DENSE>
__global__ void __launch_bounds__(192, 1) gemm_tc_kernel(const __grid_constant__ TcParams p) {
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = KC * (A_BYTES + B_BYTES);  // smem reserved per stage
  constexpr int B_OFF = KC * A_BYTES;
  constexpr int ACC_COLS = BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256));  // one accumulator buffer
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES];
  __shared__ __align__(8) uint64_t bar_empty[STAGES];
  __shared__ __align__(8) uint64_t bar_acc_full[2];
  __shared__ __align__(8) uint64_t bar_acc_empty[2];
  __shared__ uint32_t tmem_base_slot;
  constexpr int CV_IMGS = 8;                      // images one 128-pixel tile can span in the table below (hw >= 16)
  __shared__ __align__(16) float s_colvec[CV_IMGS][BN];        // bias[n] + rowvec[image, n] of the current work item

  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B atoms need 1024-B alignment
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_work = p.mt * p.nt * p.splits;
  // persistent grid (<= one CTA per SM, all resident): releasing the dependents at entry cannot starve this grid's own CTAs
  if (p.opt & 16) pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.nseg; ++i) tc::tma_prefetch_desc(&p.tmA[i][KC - 1]);
    tc::tma_prefetch_desc(&p.tmB[KC - 1]);
    for (int s = 0; s < STAGES; ++s) {
      tc::mbar_init(tc::smem_u32(&bar_full[s]), 1);
      tc::mbar_init(tc::smem_u32(&bar_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&bar_acc_full[b]), 1);
      tc::mbar_init(tc::smem_u32(&bar_acc_empty[b]), 128);
    }
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<2 * ACC_COLS>(tc::smem_u32(&tmem_base_slot));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
#ifdef MVLDM_ENABLE_PDL
  // Weights never depend on the previous kernel: pull this CTA's first weight tile (BN rows x this split's K range) into
  // L2 while the predecessor is still running.  One bulk prefetch per row, rows dealt over the lanes of the idle warp 1.
  if ((p.opt & 8) && warp == 1 && (int)blockIdx.x < num_work) {
    const int w0 = blockIdx.x;
    const int ntile = (w0 / p.mt) % p.nt, z = w0 / (p.mt * p.nt);
    // K chunk range of split z: steps are (segment, tap, channel block) in K order, KC chunks per step except at segment ends;
    // an over-estimate of a few chunks only prefetches a little more of the same rows
    const int c_begin = min(z * p.steps_per_split * KC, p.K / BK);
    const int c_end = min((z + 1) * p.steps_per_split * KC, p.K / BK);
    const uint32_t bytes = (uint32_t)(c_end - c_begin) * BK * 2;
    if (bytes)
      for (int r = lane; r < BN; r += 32) {
        const bf16* src = p.w + (int64_t)(ntile * BN + r) * p.K + (int64_t)c_begin * BK;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
      }
  }
#endif
  // everything above overlaps the previous kernel's tail; from here on we read its output
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      int it = 0;  // k-block counter across work items: smem stage = it % STAGES
      for (int w = blockIdx.x; w < num_work; w += gridDim.x) {
        const int mtile = w % p.mt, ntile = (w / p.mt) % p.nt, z = w / (p.mt * p.nt);
        const int n0 = ntile * BN;
        const int st_begin = z * p.steps_per_split;
        const int nst = min(p.num_steps - st_begin, p.steps_per_split);
        int img0, y0, x0;
        tile_origin<DENSE>(p, mtile, img0, y0, x0);
        const uint32_t a_chunk = DENSE ? (uint32_t)A_BYTES : (uint32_t)p.a_rows * (BK * 2);  // bytes per 64-channel chunk of the A box
        // locate (segment, tap, channel block) of the first step
        int s = 0, t = 0, cb = st_begin;
        while (cb >= p.seg[s].ntaps * p.seg[s].spt) {
          cb -= p.seg[s].ntaps * p.seg[s].spt;
          ++s;
        }
        t = cb / p.seg[s].spt;
        cb = (cb - t * p.seg[s].spt) * KC;
        for (int i = 0; i < nst; ++i, ++it) {
          const TcSeg& sg = p.seg[s];
          const int kc = min(KC, sg.ncblk - cb);
          const int stage = it % STAGES;
          tc::mbar_wait(tc::smem_u32(&bar_empty[stage]), ((it / STAGES) & 1) ^ 1);
          const uint32_t full = tc::smem_u32(&bar_full[stage]);
          tc::mbar_expect_tx(full, kc * (a_chunk + B_BYTES));
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          tc::tma_load_5d(sa, &p.tmA[s][kc - 1], full, 0, x0 * sg.stride + sg.dw[t], y0 * sg.stride + sg.dh[t], img0,
                          sg.coff[t] / BK + cb);
          tc::tma_load_3d(sa + B_OFF, &p.tmB[kc - 1], full, 0, n0, sg.kchunk0 + t * sg.ncblk + cb);
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
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // the whole warp runs the warp-uniform loop and the barrier waits; one elected lane issues tcgen05.mma / commit,
    // which lets ptxas emit the UTCHMMAs back to back instead of one ELECT/branch loop per instruction
    {
      constexpr uint32_t idesc = tc::umma_idesc_bf16(BM, BN, false, false);
      int it = 0, wi = 0;
      for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wi) {
        const int z = w / (p.mt * p.nt);
        const int st_begin = z * p.steps_per_split;
        const int nst = min(p.num_steps - st_begin, p.steps_per_split);
        int s = 0, t = 0, cb = st_begin;  // same walk as the producer, to know how many chunks each step carries
        while (cb >= p.seg[s].ntaps * p.seg[s].spt) {
          cb -= p.seg[s].ntaps * p.seg[s].spt;
          ++s;
        }
        t = cb / p.seg[s].spt;
        cb = (cb - t * p.seg[s].spt) * KC;
        const int ab = wi & 1;
        tc::mbar_wait(tc::smem_u32(&bar_acc_empty[ab]), ((wi >> 1) & 1) ^ 1);  // epilogue has drained this buffer
        tc::tc_fence_after();
        const uint32_t tmem_d = tmem_base + ab * ACC_COLS;
        const uint32_t a_chunk = DENSE ? (uint32_t)A_BYTES : (uint32_t)p.a_rows * (BK * 2);
        for (int i = 0; i < nst; ++i, ++it) {
          const int kc = min(KC, p.seg[s].ncblk - cb);
          const int stage = it % STAGES;
          tc::mbar_wait(tc::smem_u32(&bar_full[stage]), (it / STAGES) & 1);
          tc::tc_fence_after();
          const uint32_t sa = smem_base + stage * STAGE_BYTES;
          if (tc::elect_one()) {
            for (int c = 0; c < kc; ++c) {
              const uint64_t adesc = tc::umma_desc_k_sw128(sa + c * a_chunk);
              const uint64_t bdesc = tc::umma_desc_k_sw128(sa + B_OFF + c * B_BYTES);
#pragma unroll
              for (int k = 0; k < BK / 16; ++k)  // +32 bytes (=2 in descriptor units) per K=16 slice inside the swizzle atom
                tc::umma_ss(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (i | c | k) != 0);
            }
            tc::umma_commit(tc::smem_u32(&bar_empty[stage]));  // frees the smem slot when these MMAs retire
          }
          __syncwarp();
          cb += kc;
          if (cb == p.seg[s].ncblk) {
            cb = 0;
            if (++t == p.seg[s].ntaps) {
              t = 0;
              ++s;
            }
          }
        }
        if (tc::elect_one()) tc::umma_commit(tc::smem_u32(&bar_acc_full[ab]));
        __syncwarp();
      }
    }
  } else {
    // ================= epilogue =================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    int wi = 0;
    for (int w = blockIdx.x; w < num_work; w += gridDim.x, ++wi) {
    const int mtile = w % p.mt, ntile = (w / p.mt) % p.nt, z = w / (p.mt * p.nt);
    const int n0 = ntile * BN;
    const int ab = wi & 1;
    const uint32_t tmem_d = tmem_base + ab * ACC_COLS;
    int img0, y0, x0, di;
    tile_origin<DENSE>(p, mtile, img0, y0, x0);
    const int mtok = row_token<DENSE>(p, mtile, img0, y0, x0, row, di);
    const bool ok = mtok >= 0;
    const int m = ok ? mtok : 0;
    if (!ok) di = 0;
    const int img = img0 + di;
    // ---- while the main loop of this item runs: stage everything the epilogue needs that is not the accumulator
    const int imgs_in_tile = DENSE ? (BM + p.hw - 1) / p.hw : p.bn;
    const bool use_table = (p.opt & 1) && (!p.partial || p.counters) && (p.bias || p.rowvec) && imgs_in_tile <= CV_IMGS;
    asm volatile("bar.sync 1, 128;" ::: "memory");  // previous item's readers of s_colvec are done
    if (use_table) {
      const int et = threadIdx.x - 64;
      for (int i = et; i < imgs_in_tile * BN; i += 128) {
        const int b = i / BN, c = i - b * BN;
        float v = p.bias ? p.bias[n0 + c] : 0.f;
        if (p.rowvec && img0 + b < p.n_img) v += p.rowvec[(int64_t)(img0 + b) * p.rowvec_ld + n0 + c];
        s_colvec[b][c] = v;
      }
    }
    // residual (bf16 row of this thread): chunk c+1 is fetched while chunk c is converted and stored
    const bool use_res = ok && !p.partial && p.mode == 0 && p.residual;
    const uint4* res_row = use_res ? reinterpret_cast<const uint4*>(p.residual + (int64_t)m * p.res_ld + n0) : nullptr;
    uint4 res_next[4];
    if (use_res) {
#pragma unroll
      for (int j = 0; j < 4; ++j) res_next[j] = res_row[j];
    }
    asm volatile("bar.sync 1, 128;" ::: "memory");
    const float* cvrow = s_colvec[use_table ? di : 0];
    tc::mbar_wait(tc::smem_u32(&bar_acc_full[ab]), (wi >> 1) & 1);
    tc::tc_fence_after();
    // last item's main loop is done: let the next kernel's CTAs be scheduled (they set up barriers / TMEM / descriptors
    // and then block in griddepcontrol.wait until this grid has completed).  Triggering earlier would let them take
    // shared memory and TMEM this grid still needs.
    if (!(p.opt & 16) && w + (int)gridDim.x >= num_work) pdl_launch_dependents();
#pragma unroll 1  // rolled: the unrolled epilogue (x8 chunks x 3 modes) cost 0.5 ms per forward in code size / registers
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      uint4 res_cur[4];
      if (use_res) {
#pragma unroll
        for (int j = 0; j < 4; ++j) res_cur[j] = res_next[j];
        if (c0 + 32 < BN) {
#pragma unroll
          for (int j = 0; j < 4; ++j) res_next[j] = res_row[(c0 + 32) / 8 + j];
        }
      }
      __syncwarp();
      tc::tmem_ld32(tmem_d + ((uint32_t)(q * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
      if (ok && p.partial) {  // split-K: raw fp32 partial, reduced (+ epilogue) by splitk_reduce_kernel
        float* pp = p.partial + ((int64_t)z * p.M + m) * p.N + n0 + c0;
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
        if (p.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + n + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
        if (p.rowvec) {
          const float* rv = p.rowvec + (int64_t)img * p.rowvec_ld + n;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(rv + j);
            v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
          }
        }
      }
      if (p.mode == 0 || p.mode == 3 || p.mode == 5) {
        if (use_res) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = res_cur[j];
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
              v[j * 8 + e * 2] += f.x;
              v[j * 8 + e * 2 + 1] += f.y;
            }
          }
        }
        if (p.mode == 3) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __fdividef(v[j], 1.f + __expf(-v[j]));  // SiLU
        } else if (p.mode == 5) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_exact(v[j]);  // nn.GELU() (erf form)
        }
        bf16* op = reinterpret_cast<bf16*>(p.out) + (int64_t)m * p.ldo + n;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          st_global_v8(op + 16 * j, pack_bf16(v[j * 16], v[j * 16 + 1]), pack_bf16(v[j * 16 + 2], v[j * 16 + 3]),
                       pack_bf16(v[j * 16 + 4], v[j * 16 + 5]), pack_bf16(v[j * 16 + 6], v[j * 16 + 7]),
                       pack_bf16(v[j * 16 + 8], v[j * 16 + 9]), pack_bf16(v[j * 16 + 10], v[j * 16 + 11]),
                       pack_bf16(v[j * 16 + 12], v[j * 16 + 13]), pack_bf16(v[j * 16 + 14], v[j * 16 + 15]));
      } else if (p.mode == 1) {
        // columns [16i, 16i+8) = values, [16i+8, 16i+16) = gates of the same 8 hidden channels
        float g[16];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) g[8 * i + j] = v[16 * i + j] * gelu_exact(v[16 * i + 8 + j]);
        st_global_v8(reinterpret_cast<bf16*>(p.out) + (int64_t)m * p.ldo + n / 2, pack_bf16(g[0], g[1]), pack_bf16(g[2], g[3]),
                     pack_bf16(g[4], g[5]), pack_bf16(g[6], g[7]), pack_bf16(g[8], g[9]), pack_bf16(g[10], g[11]),
                     pack_bf16(g[12], g[13]), pack_bf16(g[14], g[15]));
      } else if (p.mode == 4) {
        float* op = reinterpret_cast<float*>(p.out) + (int64_t)m * p.ldo + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          st_global_v8(op + 8 * j, __float_as_uint(v[8 * j]), __float_as_uint(v[8 * j + 1]), __float_as_uint(v[8 * j + 2]),
                       __float_as_uint(v[8 * j + 3]), __float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]),
                       __float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]));
      } else {
        const int pix = m - img * p.hw;
        float* op = reinterpret_cast<float*>(p.out) + (int64_t)img * p.n_valid * p.hw + pix;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (n + j < p.n_valid) op[(int64_t)(n + j) * p.hw] = v[j];
      }
      }
    }
    tc::tc_fence_before();
    tc::mbar_arrive(tc::smem_u32(&bar_acc_empty[ab]));  // this thread is done reading the accumulator buffer
    if (p.counters) {
      // ---- split-K reduction fused into the GEMM: all splits of a tile are co-resident (one work item per CTA),
      // so they can meet at a global counter; each then reduces 1/splits of the tile's rows in fixed z order
      // (bit-stable) and applies the epilogue.
      const int tile = mtile + ntile * p.mt;
      const int et = threadIdx.x - 64;  // 0..127 among the epilogue threads
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {  // one cumulative gpu-scope fence publishes the 128 threads' partial rows (__threadfence() in every
                      // thread is MEMBAR.SC + an L1 invalidate each); acquire on the way out
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
        atomicAdd(&p.counters[tile], 1);
        uint32_t spins = 0;
        unsigned seen;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(&p.counters[tile]) : "memory");
          if (++spins > (1u << 26)) __trap();
        } while (seen < (unsigned)p.splits);
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const int a_rows = DENSE ? BM : p.a_rows;
      const int rows_per = (a_rows + p.splits - 1) / p.splits;
      const int r0 = min(a_rows, z * rows_per), r1 = min(a_rows, r0 + rows_per);
      constexpr int NV = BN / 8;
      const int items = (r1 - r0) * NV;
      const int64_t zstride = (int64_t)p.M * p.N;
      for (int i0 = et; i0 < items; i0 += 128 * 4) {
        // 4 independent items per thread.  Dead slots (past the end / past M) alias the thread's first item so that
        // every load below is unconditional and the 4 x 2 x splits requests are all in flight together; only the
        // final store is predicated.
        int mmv[4], nnv[4], div[4];
        bool live[4];
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * 128;
          const int ii = i < items ? i : i0;
          const int tok = row_token<DENSE>(p, mtile, img0, y0, x0, r0 + ii / NV, div[u]);
          live[u] = i < items && tok >= 0;
          mmv[u] = tok >= 0 ? tok : 0;  // dead slots read token 0 (loads stay unconditional), only the store is predicated
          if (tok < 0) div[u] = 0;
          nnv[u] = (ii % NV) * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
        }
        uint4 rres[4];
        float4 cv0[4], cv1[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          rres[u] = p.residual ? *reinterpret_cast<const uint4*>(p.residual + (int64_t)mmv[u] * p.res_ld + n0 + nnv[u])
                               : make_uint4(0u, 0u, 0u, 0u);
          if (use_table) {
            const float* cr = s_colvec[div[u]] + nnv[u];
            cv0[u] = *reinterpret_cast<const float4*>(cr);
            cv1[u] = *reinterpret_cast<const float4*>(cr + 4);
          } else {
            float t8[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              t8[j] = p.bias ? p.bias[n0 + nnv[u] + j] : 0.f;
              if (p.rowvec) t8[j] += p.rowvec[(int64_t)(mmv[u] / p.hw) * p.rowvec_ld + n0 + nnv[u] + j];
            }
            cv0[u] = make_float4(t8[0], t8[1], t8[2], t8[3]);
            cv1[u] = make_float4(t8[4], t8[5], t8[6], t8[7]);
          }
        }
#pragma unroll 4
        for (int zz = 0; zz < p.splits; ++zz) {  // fixed z order: bit-stable
          float4 a[4], b4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float4* pp = reinterpret_cast<const float4*>(p.partial + zz * zstride + (int64_t)mmv[u] * p.N + n0 + nnv[u]);
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
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wds[e]));
            v[u][2 * e] += f.x;
            v[u][2 * e + 1] += f.y;
          }
          if (live[u])
            *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(p.out) + (int64_t)mmv[u] * p.ldo + n0 + nnv[u]) =
                make_uint4(pack_bf16(v[u][0], v[u][1]), pack_bf16(v[u][2], v[u][3]), pack_bf16(v[u][4], v[u][5]),
                           pack_bf16(v[u][6], v[u][7]));
        }
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (et == 0) {  // the last split to finish re-arms the counters for the next launch
        if (atomicAdd(&p.counters[p.mt * p.nt + tile], 1) == p.splits - 1) {
          p.counters[tile] = 0;
          p.counters[p.mt * p.nt + tile] = 0;
        }
      }
    }
    }
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<2 * ACC_COLS>(tmem_base);
  }
}

// split-K second pass: out[m, n] = sum_z partial[z][m][n] (fixed order: bit-stable) + bias + rowvec + residual -> bf16
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, int splits, int M, int N,
                                                            int hw, const float* __restrict__ bias,
                                                            const float* __restrict__ rowvec, int rowvec_ld,
                                                            const bf16* __restrict__ residual, int res_ld,
                                                            bf16* __restrict__ out, int ldo) {
  pdl_wait();
  pdl_launch_dependents();
  const int nv = N / 8;
  const int64_t total = (int64_t)M * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / nv), n = (int)(i % nv) * 8;
    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int z = 0; z < splits; ++z) {
      const float4* pp = reinterpret_cast<const float4*>(partial + ((int64_t)z * M + m) * N + n);
      const float4 a = pp[0], b = pp[1];
      v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
    }
    if (bias) {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += bias[n + j];
    }
    if (rowvec) {
      const float* rv = rowvec + (int64_t)(m / hw) * rowvec_ld + n;
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] += rv[j];
    }
    if (residual) {
      const uint4 u = *reinterpret_cast<const uint4*>(residual + (int64_t)m * res_ld + n);
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
        v[2 * e] += f.x;
        v[2 * e + 1] += f.y;
      }
    }
    *reinterpret_cast<uint4*>(out + (int64_t)m * ldo + n) =
        make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
}

template <int BN, int STAGES, bool DENSE>
void launch_geom(cudaStream_t s, const TcParams& q, int grid) {
  constexpr int smem = STAGES * KC * (A_BYTES + BN * BK * 2) + 1024;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    MV_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, STAGES, DENSE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  launch_pdl(gemm_tc_kernel<BN, STAGES, DENSE>, dim3(grid), dim3(192), smem, s, q);
}

template <int BN, int STAGES>
void launch(cudaStream_t s, const TcParams& p, int splits) {
  const int num_sms = sm_count();  // SM count of the current device
  TcParams q = p;
  q.nt = p.N / BN;
  q.splits = splits;
  const int grid = std::min(q.mt * q.nt * q.splits, num_sms);
  if (p.dense) launch_geom<BN, STAGES, true>(s, q, grid);
  else launch_geom<BN, STAGES, false>(s, q, grid);
}

// ---- tile / split-K selection ---------------------------------------------------------------------
// Measured on B200 (profiles/r01_*): one SM ingests at most ~67 GB/s of TMA traffic from L2, so a GEMM here is
// bound by  (bytes the busiest SM has to pull) / 67 GB/s  long before the tensor pipe saturates.  The model
// below picks the N-tile (arithmetic intensity per SM) and the split-K factor (SMs kept busy) that minimise
//   max(load time, MMA time) + split-K reduction time.
struct TileChoice {
  int bn, splits;
};

// M-tile geometry (see TcParams): dense 128-token tiles when the output width divides 128 and an image is a whole number of
// tiles (or a tile a whole number of images); otherwise the box of (bw pixels x bh rows x bn images) that needs the fewest
// tiles, with a_rows = bw*bh*bn <= 128 a multiple of 8 (SWIZZLE_128B atoms are 8 rows: every 64-channel chunk of the box
// must start on a 1024-byte boundary).
struct TileGeom {
  int dense, bw, bh, bn, tx, ty, a_rows, mt;
};
bool tile_geometry(int n_img, int oh, int ow, TileGeom& g) {
  const int hw = oh * ow;
  if (ow <= BM && BM % ow == 0 && (hw % BM == 0 || BM % hw == 0)) {
    g.dense = 1;
    g.bw = ow;
    g.bh = std::min(oh, BM / ow);
    g.bn = BM / (g.bw * g.bh);
    g.tx = 1;
    g.ty = ceil_div(oh, g.bh);
    g.a_rows = BM;
    g.mt = ceil_div(n_img * hw, BM);
    return true;
  }
  g.dense = 0;
  g.tx = ceil_div(ow, BM);
  while (g.tx <= ow && ow % g.tx != 0) ++g.tx;  // equal column blocks
  g.bw = ow / g.tx;
  if (g.bw > BM) return false;
  long best_tiles = -1;
  for (int bh = 1; bh <= std::min(oh, BM / g.bw); ++bh)
    for (int bn = 1; bn <= BM / (g.bw * bh) && bn <= 256; ++bn) {
      const int rows = g.bw * bh * bn;
      if (rows % 8 != 0) continue;
      const long tiles = (long)g.tx * ceil_div(oh, bh) * ceil_div(n_img, bn);
      if (best_tiles < 0 || tiles < best_tiles || (tiles == best_tiles && rows > g.a_rows)) {
        best_tiles = tiles;
        g.bh = bh; g.bn = bn; g.a_rows = rows;
      }
    }
  if (best_tiles < 0) return false;
  g.ty = ceil_div(oh, g.bh);
  g.mt = (int)best_tiles;
  return true;
}

int count_steps(const mvldm_gemm_desc& d) {
  int steps = 0;
  for (int i = 0; i < d.nseg; ++i) steps += d.seg[i].ntaps * ceil_div(d.seg[i].c / BK, KC);
  return steps;
}

TileChoice pick_tiles(const mvldm_gemm_desc& d) {
  static const int kBN[5] = {256, 160, 128, 64, 32};
  static const int kSplits[12] = {1, 2, 3, 4, 5, 6, 8, 10, 12, 16, 24, 32};
  const int M = d.n_img * d.oh * d.ow, num_steps = count_steps(d);
  TileGeom geom{};
  const int mt = tile_geometry(d.n_img, d.oh, d.ow, geom) ? geom.mt : ceil_div(M, BM);
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
      const double per_sm = std::ceil(ctas / (double)sm_count());  // work items the busiest SM runs
      // measured (tools/micro/tma_ingest.cu): ~4.3 TMA boxes/us per SM whatever their size, <= ~150 GB/s per SM
      const double step_bytes = kb_per_step * (A_BYTES + bn * 128.0);
      const double t_step = std::max(std::max(2.0 / 4.3e6, step_bytes / 150e9), kb_per_step * 4.0 * (bn / 2.0) / 1.9e9);
      const double main = st_per * t_step;
      const double epi = bn * (d.mode == 1 ? 12e-9 : 6e-9);  // TMEM -> registers -> global, per item
      // persistent CTA: ramp once, items back to back (epilogue hidden behind the next main loop), last epilogue exposed
      const double t_sm = 1.5e-6 + per_sm * std::max(main, epi) + epi;
      double t_red = splits > 1 ? (splits + 1.0) * M * (double)d.n * 4.0 / 3e12 + 3e-6 : 0.0;
      // more work items than SMs: the splits of a tile are not co-resident, so the reduction is a second kernel over the
      // partials (profiles/r02_gemm_config_sweep.txt: M512 N1280 K23040 41.6 us at 6 splits = 192 items, 34.8 us at 4 = 128)
      if (splits > 1 && ctas > (double)sm_count()) t_red += 6e-6;
      const double t = t_sm + t_red;
      if (t < best_t) {
        best_t = t;
        best = TileChoice{bn, splits};
      }
    }
  }
  return best;
}

}  // namespace

constexpr size_t kCounterBytes = 2 * 4096 * sizeof(int);  // arrive/done counters live at the head of the workspace

size_t gemm_tc_workspace_bytes(const mvldm_gemm_desc& d) {
  if (count_steps(d) == 0) return 0;
  const int splits = pick_tiles(d).splits;
  return splits > 1 ? kCounterBytes + (size_t)splits * d.n_img * d.oh * d.ow * d.n * sizeof(float) : 0;
}

void gemm_tc(cudaStream_t s, const mvldm_gemm_desc& d, void* workspace, size_t workspace_bytes) {
  TcParams p{};
  const int hw = d.oh * d.ow;
  p.M = d.n_img * hw;
  p.N = d.n;
  p.hw = hw;
  p.ow = d.ow;
  MV_CHECK(d.nseg >= 1 && d.nseg <= MVLDM_MAX_SEGS, "gemm: bad segment count");
  TileGeom geom{};
  MV_CHECK(tile_geometry(d.n_img, d.oh, d.ow, geom),
           "gemm: no M-tile geometry for this output size (width " + std::to_string(d.ow) + ", height " + std::to_string(d.oh) + ")");
  // M tile = bw x bh x bn box (whole rows / whole images in the dense case)
  const int bw = geom.bw, bh = geom.bh, bn = geom.bn;
  p.dense = geom.dense; p.bw = bw; p.bh = bh; p.bn = bn; p.tx = geom.tx; p.ty = geom.ty; p.a_rows = geom.a_rows;
  p.mt = geom.mt; p.oh = d.oh; p.n_img = d.n_img;
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
    TcSeg& t = p.seg[i];
    t.ncblk = a.c / BK;
    t.ntaps = a.ntaps;
    t.stride = a.stride;
    t.spt = ceil_div(t.ncblk, KC);
    t.kchunk0 = ktot / BK;
    for (int j = 0; j < a.ntaps; ++j) {
      t.dh[j] = a.dh[j];
      t.dw[j] = a.dw[j];
      t.coff[j] = a.coff[j];
      MV_CHECK(a.coff[j] % BK == 0, "gemm: tap channel offset must be a multiple of 64");
    }
    // 5-D view (64 channels, w, h, image, 64-channel chunk): the chunk is the slowest box dimension so that one
    // box lands as [chunk][pixel][128 B] = consecutive K-major SW128 operand tiles
    const uint64_t dims[5] = {(uint64_t)BK, (uint64_t)a.sw, (uint64_t)a.sh, (uint64_t)d.n_img, (uint64_t)(a.ctot / BK)};
    const uint64_t strides[4] = {(uint64_t)a.ctot * 2, (uint64_t)a.sw * a.ctot * 2, (uint64_t)a.sh * a.sw * a.ctot * 2,
                                 (uint64_t)BK * 2};
    const uint32_t es[5] = {1, (uint32_t)a.stride, (uint32_t)a.stride, 1, 1};
    for (int kc = 1; kc <= KC; ++kc) {
      const uint32_t box[5] = {(uint32_t)BK, (uint32_t)(bw * a.stride), (uint32_t)(bh * a.stride), (uint32_t)bn, (uint32_t)kc};
      p.tmA[i][kc - 1] = make_tmap_bf16(a.ptr, 5, dims, strides, box, es);
    }
    ktot += a.c * a.ntaps;
  }
  p.nseg = d.nseg;
  MV_CHECK(ktot == d.k, "gemm: K mismatch between segments and weights");
  p.num_steps = count_steps(d);
  MV_CHECK(d.mode != 2 || d.n == 32, "gemm: NCHW head output expects N padded to 32");
  {
    const uint64_t dims[3] = {(uint64_t)BK, (uint64_t)d.n, (uint64_t)(d.k / BK)};
    const uint64_t strides[2] = {(uint64_t)d.k * 2, (uint64_t)BK * 2};
    const uint32_t es[3] = {1, 1, 1};
    for (int kc = 1; kc <= KC; ++kc) {
      const uint32_t box[3] = {(uint32_t)BK, (uint32_t)BN, (uint32_t)kc};
      p.tmB[kc - 1] = make_tmap_bf16(d.w, 3, dims, strides, box, es);
    }
  }
  int splits = tile.splits;
  if (splits > 1 && gemm_tc_workspace_bytes(d) > workspace_bytes) splits = 1;  // no scratch: plain single-pass GEMM
  p.steps_per_split = ceil_div(p.num_steps, splits);
  splits = ceil_div(p.num_steps, p.steps_per_split);
  p.partial = splits > 1 ? reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + kCounterBytes) : nullptr;
  // fused reduction needs every split of a tile resident at once: one work item per CTA, grid <= #SMs
  const int work = p.mt * (d.n / BN) * splits;
  // (one CTA per SM: shared memory; the kernel is launched with min(work, #SMs) CTAs, so work <= #SMs means every split
  // of every tile has its own resident CTA on an otherwise idle device; more work takes the two-pass reduction)
  const bool fused = splits > 1 && work <= sm_count() && p.mt * (d.n / BN) <= 4096;
  p.counters = fused ? reinterpret_cast<int*>(workspace) : nullptr;
  {
    static const int opt = [] {
      const char* e = getenv("MVLDM_GEMM_OPT");
      return e ? atoi(e) : 7;
    }();
    p.opt = opt;
  }
  p.w = reinterpret_cast<const bf16*>(d.w);
  p.K = d.k;
  p.bias = d.bias;
  p.rowvec = d.rowvec;
  p.rowvec_ld = d.rowvec_ld;
  p.residual = reinterpret_cast<const bf16*>(d.residual);
  p.res_ld = d.res_ld;
  p.mode = d.mode;
  p.out = d.out;
  p.ldo = d.ldo;
  p.n_valid = d.n_valid;
  MV_CHECK(d.mode >= 0 && d.mode <= 5, "gemm: bad output mode");
  if (d.mode == 0 || d.mode == 3 || d.mode == 4 || d.mode == 5) MV_CHECK(d.ldo % 8 == 0 && (!d.residual || d.res_ld % 8 == 0), "gemm: row pitch must be a multiple of 8");
  if (BN == 256) launch<256, 2>(s, p, splits);        // 2 x 96 KB
  else if (BN == 160) launch<160, 3>(s, p, splits);   // 3 x 72 KB
  else if (BN == 128) launch<128, 3>(s, p, splits);   // 3 x 64 KB
  else if (BN == 64) launch<64, 4>(s, p, splits);     // 4 x 48 KB
  else launch<32, 5>(s, p, splits);                   // 5 x 40 KB
  if (splits > 1 && !fused) {
    const int64_t total = (int64_t)p.M * (p.N / 8);
    const int blocks = (int)std::min<int64_t>((total + 255) / 256, 148 * 8);
    launch_pdl(splitk_reduce_kernel, dim3(blocks), dim3(256), 0, s, (const float*)p.partial, splits, p.M, p.N, p.hw, p.bias,
               p.rowvec, p.rowvec_ld, p.residual, p.res_ld, reinterpret_cast<bf16*>(p.out), p.ldo);
  }
}

}  // namespace mvldm
