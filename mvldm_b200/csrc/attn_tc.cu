// Flash-style self-attention on tcgen05 / TMEM / TMA for sm_100a: the joint multi-view attention (one
// sequence = all V*h*w tokens of a scene) and the per-view attention of BasicTransformerBlock3D
// (reference src/model/denoiser/mvdream/attention.py:174-205,362-368).  The N x N score matrix the reference
// materialises in fp32 (2.1 GB per level-0 block at 8 views) never leaves the SM.
//
// One CTA = 128 queries of one (batch, head).  Per 128-key (64 for head_dim_pad 192) tile:
//   S = Q K^T          tcgen05.mma  SS  (Q, K in smem via TMA, K-major SW128)      -> TMEM fp32
//   softmax            4 warps, one query row per thread: tcgen05.ld S in 32-column chunks, exp2 against a lazily
//                      moved running reference, row sum in fp32, P -> bf16 into its own TMEM columns (tcgen05.st)
//   O += P V           tcgen05.mma  TS  (P from TMEM, V from smem as an MN-major SW128 operand) -> TMEM fp32
// Scores are fp32 and the softmax is fp32 exactly as ATTN_PRECISION=fp32 (attention.py:185-203); P is
// rounded to bf16 before P.V as autocast does at :203.
//
// Warp roles (192 threads): warp 0 TMA producer, warp 1 TMEM owner + MMA issuer, warps 2..5 softmax/epilogue.
// The running maximum is only moved when it grows by more than 2^8 (exact: any reference point is a valid
// softmax shift; P <= 256 stays well inside bf16 range), so the O accumulator is almost never touched.
#include <cstdlib>

#include "common.cuh"
#include "tc_common.cuh"

namespace mvldm {
namespace {

constexpr int BM = 128;

// optional in-kernel timeline (tools/attn_trace.py): CTA (0,0) stamps clock64() at the pipeline hand-offs
__device__ long long g_attn_trace[8 * 512];
__device__ __forceinline__ void trace(bool on, int slot, int j) {
  if (on && j < 512) g_attn_trace[slot * 512 + j] = clock64();
}

struct AttnParams {
  CUtensorMap tmQ;   // box [64, 128, 1] over (cols, seq, batch)
  CUtensorMap tmKV;  // box [64, BN, 1]
  CUtensorMap tmKV2; // 4-D (cols, keys, {K,V}, batch), box [64, 128, 2, 1]: one TMA instruction lands a K tile and its V tile
  bf16* out;
  int seq;             // queries per batch (rows of the Q source and of the output)
  int seq_kv;          // keys per batch (rows of the K/V source); == seq for plain self-attention
  int q_col0, k_col0, v_col0;  // first column of head 0 inside the Q source / the K/V source
  int heads, d, dpad;
  float scale_log2;  // d^-1/2 * log2(e)
  int trace;
  const bf16* q_ptr; // Q rows for the kernels that stage Q through registers into TMEM
  int ld_q;
  float2* stats;     // optional [batches*seq, heads] (reference max * scale_log2, row sum against it): partial-softmax state for
                     // attention_merge when the keys of one softmax are spread over several launches
  int skip_pad;      // MMAs cover only ceil(d/16) K-slices / ceil((d+1)/16)*16 output columns (MVLDM_ATTN_SKIP_PAD=0: all)
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// packed fp32x2 arithmetic (sm_100: FFMA2): one issue slot for two lanes of the softmax's scale-and-shift
__device__ __forceinline__ uint64_t pack2(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void ffma2(float& o0, float& o1, uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(o0), "=f"(o1) : "l"(d));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// DPAD: padded head dim (64/128/192); BN: keys per tile; ST: K/V ring stages; OCC: CTAs per SM the budget allows
template <int DPAD, int BN, int ST, int OCC>
__global__ void __launch_bounds__(192, OCC) attn_tc_kernel(const __grid_constant__ AttnParams p) {
  constexpr int NC = DPAD / 64;                // 64-column chunks per row
  constexpr int Q_BYTES = NC * BM * 128;
  constexpr int KV_CHUNK = BN * 128;           // one 64-column chunk of a K or V tile
  constexpr int STAGE_BYTES = 2 * NC * KV_CHUNK;
  // TMEM columns: S (fp32 scores) | P (bf16 pairs) | O (fp32 accumulator).  For DPAD = 64 this is exactly 256
  // columns, so two CTAs share an SM: while one CTA's softmax warps wait for the tensor pipe the other's run,
  // which keeps the MUFU (exp2) pipe - the real bound of this kernel - busy.
  constexpr uint32_t S_COL = 0, P_COL = BN, O_COL = BN + BN / 2;
  constexpr int TMEM_COLS = (BN + BN / 2 + DPAD) <= 256 ? 256 : 512;
  static_assert(BN + BN / 2 + DPAD <= 512, "TMEM budget");
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_kv_full[ST], bar_kv_empty[ST], bar_s, bar_p, bar_o;
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = smem_base;
  const uint32_t kv_smem = smem_base + Q_BYTES;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int batch = blockIdx.y / p.heads, head = blockIdx.y % p.heads;
  const int T = (p.seq_kv + BN - 1) / BN;
  const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&p.tmQ);
    tc::tma_prefetch_desc(&p.tmKV);
    tc::mbar_init(tc::smem_u32(&bar_q), 1);
    for (int s = 0; s < ST; ++s) {
      tc::mbar_init(tc::smem_u32(&bar_kv_full[s]), 1);
      tc::mbar_init(tc::smem_u32(&bar_kv_empty[s]), 1);
    }
    tc::mbar_init(tc::smem_u32(&bar_s), 1);
    tc::mbar_init(tc::smem_u32(&bar_p), 128);
    tc::mbar_init(tc::smem_u32(&bar_o), 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<TMEM_COLS>(tc::smem_u32(&tmem_slot));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const uint32_t bq = tc::smem_u32(&bar_q);
      tc::mbar_expect_tx(bq, Q_BYTES);
      for (int c = 0; c < NC; ++c) tc::tma_load_3d(q_smem + c * BM * 128, &p.tmQ, bq, p.q_col0 + head * DPAD + c * 64, q0, batch);
      const int kcol = p.k_col0 + head * DPAD, vcol = p.v_col0 + head * DPAD;
      for (int j = 0; j < T; ++j) {
        const int stage = j % ST;
        tc::mbar_wait(tc::smem_u32(&bar_kv_empty[stage]), ((j / ST) & 1) ^ 1);
        const uint32_t full = tc::smem_u32(&bar_kv_full[stage]);
        tc::mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t ks = kv_smem + stage * STAGE_BYTES, vs = ks + NC * KV_CHUNK;
        for (int c = 0; c < NC; ++c) tc::tma_load_3d(ks + c * KV_CHUNK, &p.tmKV, full, kcol + c * 64, j * BN, batch);
        for (int c = 0; c < NC; ++c) tc::tma_load_3d(vs + c * KV_CHUNK, &p.tmKV, full, vcol + c * 64, j * BN, batch);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = tc::umma_idesc_bf16(BM, BN, false, false);
      // the pad columns of a head are zeros: skip their K-slices in Q K^T and their output columns in P V
      const int kq = p.skip_pad ? (p.d + 15) / 16 : DPAD / 16;
      const uint32_t idesc_pv = tc::umma_idesc_bf16(BM, p.skip_pad ? (p.d + 15) / 16 * 16 : DPAD, false, true);
      auto issue_qk = [&](int j) {
        const int stage = j % ST;
        tc::mbar_wait(tc::smem_u32(&bar_kv_full[stage]), (j / ST) & 1);
        tc::tc_fence_after();
        const uint32_t ks = kv_smem + stage * STAGE_BYTES;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const uint64_t qd = tc::umma_desc_k_sw128(q_smem + c * BM * 128);
          const uint64_t kd = tc::umma_desc_k_sw128(ks + c * KV_CHUNK);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (c * 4 + k < kq) tc::umma_ss(tmem + S_COL, qd + 2 * k, kd + 2 * k, idesc_qk, (c | k) != 0);
        }
        tc::umma_commit(tc::smem_u32(&bar_s));
      };
      tc::mbar_wait(tc::smem_u32(&bar_q), 0);
      issue_qk(0);
      for (int j = 0; j < T; ++j) {
        tc::mbar_wait(tc::smem_u32(&bar_p), j & 1);  // softmax has read S(j) and written P(j)
        tc::tc_fence_after();
        trace(tr, 3, j);
        const int stage = j % ST;
        const uint32_t vs = kv_smem + stage * STAGE_BYTES + NC * KV_CHUNK;
        const uint64_t vd = tc::umma_desc_mn_sw128(vs, KV_CHUNK);
#pragma unroll
        for (int kk = 0; kk < BN / 16; ++kk)  // 16 keys per MMA: 8 packed-bf16 TMEM columns of P, 16 V rows (2 KB)
          tc::umma_ts(tmem + O_COL, tmem + P_COL + kk * 8, vd + kk * 128, idesc_pv, (j | kk) != 0);
        tc::umma_commit(tc::smem_u32(&bar_kv_empty[stage]));
        tc::umma_commit(tc::smem_u32(&bar_o));
        trace(tr, 4, j);
        if (j + 1 < T) issue_qk(j + 1);  // S is free again (in-order tensor pipe: after P V(j))
        trace(tr, 5, j);
      }
    }
    __syncwarp();
  } else {
    // ================= softmax + epilogue: one query row per thread =================
    const int quarter = warp & 3;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int row = q0 + quarter * 32 + lane;
    const float sc = p.scale_log2;
    float m_ref = -INFINITY, l_sum = 0.f;
    for (int j = 0; j < T; ++j) {
      trace(tr && threadIdx.x == 64, 0, j);
      tc::mbar_wait(tc::smem_u32(&bar_s), j & 1);
      tc::tc_fence_after();
      trace(tr && threadIdx.x == 64, 1, j);
      const int valid = (j == T - 1) ? p.seq_kv - j * BN : BN;  // ragged tail: keys past the sequence end do not exist
      // P is computed against the running reference m_ref from earlier tiles, so no separate max pass sits in front
      // of the exponentials.  If this tile raises the row maximum by more than 2^8 (always on the first tile) the
      // reference is moved, O and l are rescaled, and the tile is redone - rare after the first few tiles.
#pragma unroll 1
      for (int pass = 0; pass < 2; ++pass) {
        const float mb = m_ref * sc;
        float mt = -INFINITY, l0 = 0.f, l1 = 0.f;
        // one 32-column chunk: mask the ragged tail, exp2 against the reference, row sum, pack to bf16, hand to TMEM
        auto chunk = [&](uint32_t(&r)[32], int c) {
          if (valid < BN) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (c * 32 + i >= valid) r[i] = 0xff800000u;  // -inf
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float s0 = __uint_as_float(r[i]), s1 = __uint_as_float(r[i + 1]);
            mt = fmaxf(mt, fmaxf(s0, s1));
            const float p0 = ex2(fmaf(s0, sc, -mb)), p1 = ex2(fmaf(s1, sc, -mb));
            l0 += p0;
            l1 += p1;
            pk[i / 2] = pack_bf16(p0, p1);
          }
          tc::tmem_st16(tmem + lane_addr + P_COL + c * 16, pk);
        };
        // software pipeline over two register buffers: the TMEM load of chunk c+1 is in flight while chunk c is
        // exponentiated (tcgen05.wait::ld is issued before the next load, so it only covers the current chunk)
        uint32_t ra[32], rb[32];
        tc::tmem_ld32(tmem + lane_addr + S_COL, ra);
#pragma unroll 1
        for (int c = 0; c < BN / 32; c += 2) {
          tc::tmem_ld_wait();
          tc::tmem_ld32(tmem + lane_addr + S_COL + (c + 1) * 32, rb);
          chunk(ra, c);
          tc::tmem_ld_wait();
          if (c + 2 < BN / 32) tc::tmem_ld32(tmem + lane_addr + S_COL + (c + 2) * 32, ra);
          chunk(rb, c + 1);
        }
        const bool grow = (mt - m_ref) * sc > 8.f;  // true on the first tile (m_ref = -inf)
        if (!__any_sync(0xffffffffu, grow)) {
          l_sum += l0 + l1;
          break;
        }
        // move the reference (per row; rows that did not grow keep theirs and rescale by 1)
        const float m_new = grow ? mt : m_ref;
        if (j > 0) {
          const float f = ex2((m_ref - m_new) * sc);
          tc::mbar_wait(tc::smem_u32(&bar_o), (j - 1) & 1);  // P V of the previous tile has landed in O
          tc::tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < DPAD / 32; ++c) {
            uint32_t o[32];
            tc::tmem_ld32(tmem + lane_addr + O_COL + c * 32, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tc::tmem_st16(tmem + lane_addr + O_COL + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
            tc::tmem_st16(tmem + lane_addr + O_COL + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
          }
          l_sum *= f;
        }
        m_ref = m_new;  // second pass recomputes P(j) against the new reference (it cannot grow again)
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(tc::smem_u32(&bar_p));
      trace(tr && threadIdx.x == 64, 2, j);
    }
    // ---- epilogue: O / l -> bf16, head-padded row
    tc::mbar_wait(tc::smem_u32(&bar_o), (T - 1) & 1);
    tc::tc_fence_after();
    pdl_launch_dependents();
    const float inv = 1.f / l_sum;
    if (p.stats && row < p.seq) p.stats[((int64_t)batch * p.seq + row) * p.heads + head] = make_float2(m_ref * sc, l_sum);
    bf16* dst = p.out + ((int64_t)batch * p.seq + row) * (p.heads * DPAD) + head * DPAD;
#pragma unroll 1
    for (int c = 0; c < DPAD / 16; ++c) {  // rolled, 16 columns at a time: keeps the kernel at two CTAs per SM
      uint32_t o[16];
      __syncwarp();
      tc::tmem_ld16(tmem + lane_addr + O_COL + c * 16, o);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)  // pad columns (V carries a ones column there) are written as zeros
        if (c * 16 + i >= p.d) o[i] = 0u;
      if (row < p.seq) {
        uint4* op = reinterpret_cast<uint4*>(dst + c * 16);
#pragma unroll
        for (int v = 0; v < 2; ++v)
          op[v] = make_uint4(pack_bf16(__uint_as_float(o[v * 8]) * inv, __uint_as_float(o[v * 8 + 1]) * inv),
                             pack_bf16(__uint_as_float(o[v * 8 + 2]) * inv, __uint_as_float(o[v * 8 + 3]) * inv),
                             pack_bf16(__uint_as_float(o[v * 8 + 4]) * inv, __uint_as_float(o[v * 8 + 5]) * inv),
                             pack_bf16(__uint_as_float(o[v * 8 + 6]) * inv, __uint_as_float(o[v * 8 + 7]) * inv));
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<TMEM_COLS>(tmem);
  }
}

// ---------------------------------------------------------------------------------------------------------
// head_dim_pad = 64 (the 32x32 level: 88 % of the attention FLOPs).  Same algorithm, tuned structure:
//  * 64-key tiles with S and P double-buffered in TMEM (2x64 + 2x32 + 64 = 256 columns), so the tensor pipe
//    works on Q K^T (j+2) and P V (j) while the softmax warps are on tile j+1, and two CTAs still share an SM;
//  * K and V arrive in 128-key TMA boxes (one box feeds two tiles: the per-SM TMA box rate is the scarce resource);
//  * the softmax loop is issue-bound, so it is stripped to scale-and-shift (packed FFMA2) + ex2 + bf16 pack per
//    element: the row sum is produced by the tensor core through a ones column in V's padding (column d).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_poly(float x) {
  // 2^x for x <= ~8: x = n + f, f in [-0.5, 0.5]; 2^f by a cubic minimax; scale by 2^n through the exponent field
  x = fmaxf(x, -120.f);
  const float t = x + 12582912.f;            // 1.5 * 2^23: rounds x to the nearest integer in the low mantissa bits
  const float n = t - 12582912.f;
  const float f = x - n;
  float pl = fmaf(f, 0.05517167f, 0.24261113f);   // minimax in relative error on [-0.5, 0.5]: 7.5e-5 (bf16 half-ulp: 2e-3)
  pl = fmaf(pl, f, 0.69326097f);
  pl = fmaf(pl, f, 0.99992806f);
  return __int_as_float(__float_as_int(pl) + (__float_as_int(t) << 23));
}

#ifndef MVLDM_POLY_EVERY
#define MVLDM_POLY_EVERY 2
#endif
// One exponential in 2*POLY_EVERY goes to the FMA pipe (ex2_poly) instead of the MUFU pipe.  With Q in TMEM the kernel sits
// on the MUFU limit (two CTAs x 64 ex2 per thread per tile = 1024 of a 1134-cycle tile period); moving 1/4 of them
// measured 205.9 -> 189.8 us (1 scene, 8192 tokens) and 1346 -> 1262 us (8 scenes); 1/2 overloads the FMA pipe (207.6 us).
constexpr int POLY_EVERY = MVLDM_POLY_EVERY;

template <int ST>
__global__ void __launch_bounds__(192, 2) attn64q_kernel(const __grid_constant__ AttnParams p) {
  constexpr int BT = 64;                       // keys per compute tile
  constexpr int Q_BYTES = BM * 128;            // 128 queries x 64 cols bf16
  constexpr int KV_TILE = 128 * 128;           // 128 keys x 64 cols bf16 (one TMA box)
  constexpr int STAGE_BYTES = 2 * KV_TILE;     // K box + V box
  // TMEM: Q (bf16 pairs, the A operand of Q K^T) | S0, S1 (P(j) is written over the first half of S(j)) | O = 224 columns
  constexpr uint32_t Q_COL = 0, S_COL = 32, O_COL = 160;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_q, bar_kv_full[ST], bar_kv_empty[ST], bar_s[2], bar_p[2], bar_o, bar_done;  // bar_q: Q is in TMEM
  __shared__ uint32_t tmem_slot;

  const uint32_t smem_base = (tc::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t kv_smem = smem_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * BM;
  const int batch = blockIdx.y / p.heads, head = blockIdx.y % p.heads;
  const int T = (p.seq_kv + BT - 1) / BT;      // 64-key tiles
  const int NS = (T + 1) / 2;                  // 128-key stages
  const bool tr = p.trace && blockIdx.x == 0 && blockIdx.y == 0;

  if (warp == 0 && lane == 0) {
    tc::tma_prefetch_desc(&p.tmKV2);
    tc::mbar_init(tc::smem_u32(&bar_q), 128);
    for (int s = 0; s < ST; ++s) {
      tc::mbar_init(tc::smem_u32(&bar_kv_full[s]), 1);
      tc::mbar_init(tc::smem_u32(&bar_kv_empty[s]), 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(tc::smem_u32(&bar_s[b]), 1);
      tc::mbar_init(tc::smem_u32(&bar_p[b]), 128);
    }
    tc::mbar_init(tc::smem_u32(&bar_o), 1);
    tc::mbar_init(tc::smem_u32(&bar_done), 1);
    tc::mbar_fence_init();
  }
  if (warp == 1) tc::tmem_alloc<256>(tc::smem_u32(&tmem_slot));
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      const int kcol = p.k_col0 + head * 64;
      for (int s = 0; s < NS; ++s) {
        const int stage = s % ST;
        tc::mbar_wait(tc::smem_u32(&bar_kv_empty[stage]), ((s / ST) & 1) ^ 1);
        const uint32_t full = tc::smem_u32(&bar_kv_full[stage]);
        tc::mbar_expect_tx(full, STAGE_BYTES);
        const uint32_t ks = kv_smem + stage * STAGE_BYTES;
        tc::tma_load_4d(ks, &p.tmKV2, full, kcol, s * 128, 0, batch);  // K tile, then V tile (the TMA box rate is scarce)
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The whole warp runs the (warp-uniform) control flow and barrier waits; one elected lane issues the tcgen05
    // instructions.  Keeping the loop convergent lets ptxas keep descriptors in uniform registers and predicate
    // UTCHMMA directly instead of wrapping every issue in an ELECT/branch loop.
    {
      // The head's pad columns are zeros (Q, K) or unused (V beyond the ones column): the tensor pipe, which two CTAs
      // per SM keep ~90 % busy, skips them.  d = 40: 3 of the 4 K=16 slices of Q K^T and N = 48 of 64 columns of P V.
      constexpr uint32_t idesc_qk = tc::umma_idesc_bf16(BM, BT, false, false);
      const int kq = p.skip_pad ? (p.d + 15) / 16 : 4;
      const uint32_t idesc_pv = tc::umma_idesc_bf16(BM, p.skip_pad ? (p.d + 1 + 15) / 16 * 16 : 64, false, true);
      auto issue_qk = [&](int j) {
        const int s = j >> 1, stage = s % ST;
        if ((j & 1) == 0) {  // first tile of a stage: its K/V box must have landed
          tc::mbar_wait(tc::smem_u32(&bar_kv_full[stage]), (s / ST) & 1);
          tc::tc_fence_after();
        }
        const uint64_t kd = tc::umma_desc_k_sw128(kv_smem + stage * STAGE_BYTES + (j & 1) * (BT * 128));
        if (tc::elect_one()) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (k < kq) tc::umma_ts(tmem + S_COL + (j & 1) * BT, tmem + Q_COL + k * 8, kd + 2 * k, idesc_qk, k != 0);
          tc::umma_commit(tc::smem_u32(&bar_s[j & 1]));
        }
        __syncwarp();
      };
      tc::mbar_wait(tc::smem_u32(&bar_q), 0);
      tc::tc_fence_after();
      issue_qk(0);
      if (T > 1) issue_qk(1);
      for (int j = 0; j < T; ++j) {
        tc::mbar_wait(tc::smem_u32(&bar_p[j & 1]), (j >> 1) & 1);  // softmax has read S(j) and written P(j)
        tc::tc_fence_after();
        trace(tr && lane == 0, 3, j);
        const int s = j >> 1, stage = s % ST;
        const uint32_t vs = kv_smem + stage * STAGE_BYTES + KV_TILE + (j & 1) * (BT * 128);
        const uint64_t vd = tc::umma_desc_mn_sw128(vs, KV_TILE);
        if (tc::elect_one()) {
#pragma unroll
          for (int kk = 0; kk < BT / 16; ++kk)  // 16 keys per MMA: 8 packed-bf16 TMEM columns of P, 16 V rows (2 KB)
            tc::umma_ts(tmem + O_COL, tmem + S_COL + (j & 1) * BT + kk * 8, vd + kk * 128, idesc_pv, (j | kk) != 0);
          tc::umma_commit(tc::smem_u32(&bar_o));
          if ((j & 1) == 1 || j == T - 1) tc::umma_commit(tc::smem_u32(&bar_kv_empty[stage]));  // stage fully consumed
          // The epilogue needs "every P V has landed".  With S double-buffered a softmax warp can be two bar_o phases
          // ahead of the tensor pipe, where a parity wait on bar_o is ambiguous, hence a dedicated single-phase barrier.
          if (j == T - 1) tc::umma_commit(tc::smem_u32(&bar_done));
        }
        __syncwarp();
        trace(tr && lane == 0, 4, j);
        if (j + 2 < T) issue_qk(j + 2);  // into the S buffer softmax(j) has just released
        trace(tr && lane == 0, 5, j);
      }
    }
    __syncwarp();
  } else {
    // ================= softmax + epilogue: one query row per thread =================
    const int quarter = warp & 3;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const int row = q0 + quarter * 32 + lane;
    const float sc = p.scale_log2;
    float m_ref = -INFINITY;
    {  // this thread's query row -> TMEM: Q K^T then takes its A operand from TMEM (an SS MMA re-reads the 128-row Q
       // slice from shared memory for every 64-key tile and measured 2x the issue time of the equally sized TS-form P V)
      uint32_t qa[16], qb[16];
      const uint4* qrow = reinterpret_cast<const uint4*>(p.q_ptr + ((int64_t)batch * p.seq + min(row, p.seq - 1)) * p.ld_q +
                                                         p.q_col0 + head * 64);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = row < p.seq ? qrow[i] : make_uint4(0u, 0u, 0u, 0u);
        qa[4 * i] = u.x; qa[4 * i + 1] = u.y; qa[4 * i + 2] = u.z; qa[4 * i + 3] = u.w;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 u = row < p.seq ? qrow[4 + i] : make_uint4(0u, 0u, 0u, 0u);
        qb[4 * i] = u.x; qb[4 * i + 1] = u.y; qb[4 * i + 2] = u.z; qb[4 * i + 3] = u.w;
      }
      tc::tmem_st16(tmem + lane_addr + Q_COL, qa);
      tc::tmem_st16(tmem + lane_addr + Q_COL + 16, qb);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(tc::smem_u32(&bar_q));
    }
    for (int j = 0; j < T; ++j) {
      const uint32_t s_tm = tmem + lane_addr + S_COL + (j & 1) * BT;
      const uint32_t p_tm = s_tm;  // P(j) overwrites the first 32 columns of S(j): S is in registers by then
      trace(tr && threadIdx.x == 64, 0, j);
      tc::mbar_wait(tc::smem_u32(&bar_s[j & 1]), (j >> 1) & 1);
      tc::tc_fence_after();
      trace(tr && threadIdx.x == 64, 1, j);
      const int valid = (j == T - 1) ? p.seq_kv - j * BT : BT;  // ragged tail: keys past the sequence end do not exist
      uint32_t ra[32], rb[32];
      tc::tmem_ld32(s_tm, ra);
      tc::tmem_ld32(s_tm + 32, rb);
      tc::tmem_ld_wait();
      if (valid < BT) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= valid) ra[i] = 0xff800000u;       // -inf
          if (32 + i >= valid) rb[i] = 0xff800000u;
        }
      }
      // P = exp2(s * sc - m_ref * sc) against the reference of the EARLIER tiles, with the tile maximum tracked on the
      // side (independent instruction stream, off the critical path).  Only if this tile would have pushed P beyond 2^8
      // (always on the first tile) is the reference moved, O rescaled and P recomputed from the registers.
      // The row sum is not accumulated here: V carries 1.0 in its first pad column, so P V delivers sum_j P_ij (of the
      // bf16-rounded P the tensor core actually uses) in column d of O.
      const uint32_t scb = __float_as_uint(sc);
      const uint64_t sc2 = pack2(scb, scb);
      float mt0 = -INFINITY, mt1 = -INFINITY;
      auto half = [&](uint32_t(&r)[32], uint32_t dst, uint64_t nmb2, bool track) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float x0, x1;
          if (track) {
            mt0 = fmaxf(mt0, __uint_as_float(r[i]));
            mt1 = fmaxf(mt1, __uint_as_float(r[i + 1]));
          }
          ffma2(x0, x1, pack2(r[i], r[i + 1]), sc2, nmb2);
          // the MUFU pipe (16 ex2/clk/SM) is busy here: every POLY_EVERY-th pair sends one exponential to the FMA pipe
          const float e1 = ((i / 2) % POLY_EVERY == POLY_EVERY - 1) ? ex2_poly(x1) : ex2(x1);
          pk[i / 2] = pack_bf16(ex2(x0), e1);
        }
        tc::tmem_st16(dst, pk);
      };
      {
        const uint32_t nmb = __float_as_uint(-m_ref * sc);
        const uint64_t nmb2 = pack2(nmb, nmb);
        half(ra, p_tm, nmb2, true);
        half(rb, p_tm + 16, nmb2, true);
      }
      const float mt = fmaxf(mt0, mt1);
      const bool grow = (mt - m_ref) * sc > 8.f;  // also true on the first tile (m_ref = -inf)
      if (__any_sync(0xffffffffu, grow)) {
        const float m_new = grow ? mt : m_ref;
        if (j > 0) {
          const float f = ex2((m_ref - m_new) * sc);
          tc::mbar_wait(tc::smem_u32(&bar_o), (j - 1) & 1);  // P V of the previous tile has landed in O
          tc::tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t o[32];
            tc::tmem_ld32(tmem + lane_addr + O_COL + c * 32, o);
            tc::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tc::tmem_st16(tmem + lane_addr + O_COL + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[0]));
            tc::tmem_st16(tmem + lane_addr + O_COL + c * 32 + 16, *reinterpret_cast<uint32_t(*)[16]>(&o[16]));
          }
        }
        m_ref = m_new;
        const uint32_t nmb = __float_as_uint(-m_ref * sc);
        const uint64_t nmb2 = pack2(nmb, nmb);
        half(ra, p_tm, nmb2, false);   // recompute P(j) against the moved reference (S is still in registers)
        half(rb, p_tm + 16, nmb2, false);
      }
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::mbar_arrive(tc::smem_u32(&bar_p[j & 1]));
      trace(tr && threadIdx.x == 64, 2, j);
    }
    // ---- epilogue: O / l -> bf16, head-padded row
    tc::mbar_wait(tc::smem_u32(&bar_done), 0);
    tc::tc_fence_after();
    pdl_launch_dependents();
    float inv;
    {  // row sum = column d of O (the ones column of V)
      uint32_t o[16];
      tc::tmem_ld16(tmem + lane_addr + O_COL + (p.d / 16) * 16, o);
      tc::tmem_ld_wait();
      float l = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) l = (i == (p.d & 15)) ? __uint_as_float(o[i]) : l;
      inv = 1.f / l;
      if (p.stats && row < p.seq) p.stats[((int64_t)batch * p.seq + row) * p.heads + head] = make_float2(m_ref * sc, l);
    }
    bf16* dst = p.out + ((int64_t)batch * p.seq + row) * (p.heads * 64) + head * 64;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[16];
      __syncwarp();
      tc::tmem_ld16(tmem + lane_addr + O_COL + c * 16, o);
      tc::tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i)  // pad columns (including the row-sum column) are written as zeros
        if (c * 16 + i >= p.d) o[i] = 0u;
      if (row < p.seq) {
        uint4* op = reinterpret_cast<uint4*>(dst + c * 16);
#pragma unroll
        for (int v = 0; v < 2; ++v)
          op[v] = make_uint4(pack_bf16(__uint_as_float(o[v * 8]) * inv, __uint_as_float(o[v * 8 + 1]) * inv),
                             pack_bf16(__uint_as_float(o[v * 8 + 2]) * inv, __uint_as_float(o[v * 8 + 3]) * inv),
                             pack_bf16(__uint_as_float(o[v * 8 + 4]) * inv, __uint_as_float(o[v * 8 + 5]) * inv),
                             pack_bf16(__uint_as_float(o[v * 8 + 6]) * inv, __uint_as_float(o[v * 8 + 7]) * inv));
      }
    }
    tc::tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc::tc_fence_after();
    tc::tmem_dealloc<256>(tmem);
  }
}

struct AttnSrc {  // where the operands live: Q rows [batches*seq_q, ld_q], K/V rows [batches*seq_kv, ld_kv]
  const bf16* q;
  int ld_q, q_col0;
  const bf16* kv;
  int ld_kv, k_col0, v_col0;
  int seq_q, seq_kv;
};

void fill_params(AttnParams& p, const AttnSrc& a, bf16* out, int batches, int heads, int d, int dpad, int bn_rows) {
  const uint32_t es[3] = {1, 1, 1};
  {
    const uint64_t dims[3] = {(uint64_t)a.ld_q, (uint64_t)a.seq_q, (uint64_t)batches};
    const uint64_t strides[2] = {(uint64_t)a.ld_q * 2, (uint64_t)a.seq_q * a.ld_q * 2};
    const uint32_t box[3] = {64, BM, 1};
    p.tmQ = make_tmap_bf16(a.q, 3, dims, strides, box, es);
  }
  {
    const uint64_t dims[3] = {(uint64_t)a.ld_kv, (uint64_t)a.seq_kv, (uint64_t)batches};
    const uint64_t strides[2] = {(uint64_t)a.ld_kv * 2, (uint64_t)a.seq_kv * a.ld_kv * 2};
    const uint32_t box[3] = {64, (uint32_t)bn_rows, 1};
    p.tmKV = make_tmap_bf16(a.kv, 3, dims, strides, box, es);
  }
  {
    const uint64_t dims[4] = {(uint64_t)a.ld_kv, (uint64_t)a.seq_kv, 2, (uint64_t)batches};
    const uint64_t strides[3] = {(uint64_t)a.ld_kv * 2, (uint64_t)(a.v_col0 - a.k_col0) * 2, (uint64_t)a.seq_kv * a.ld_kv * 2};
    const uint32_t box[4] = {64, 128, 2, 1};
    const uint32_t es4[4] = {1, 1, 1, 1};
    if (a.v_col0 > a.k_col0 && (a.v_col0 - a.k_col0) % 8 == 0) p.tmKV2 = make_tmap_bf16(a.kv, 4, dims, strides, box, es4);
  }
  p.out = out;
  p.q_ptr = a.q;
  p.ld_q = a.ld_q;
  p.seq = a.seq_q;
  p.seq_kv = a.seq_kv;
  p.q_col0 = a.q_col0;
  p.k_col0 = a.k_col0;
  p.v_col0 = a.v_col0;
  p.heads = heads;
  p.d = d;
  p.dpad = dpad;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)d);
  p.trace = getenv("MVLDM_ATTN_TRACE") != nullptr;
  static const bool skip = [] {
    const char* e = getenv("MVLDM_ATTN_SKIP_PAD");
    return !e || atoi(e) != 0;
  }();
  p.skip_pad = skip ? 1 : 0;
}

template <int DPAD, int BN, int ST, int OCC>
void launch(cudaStream_t s, const AttnSrc& a, bf16* out, float2* stats, int batches, int heads, int d) {
  AttnParams p{};
  fill_params(p, a, out, batches, heads, d, DPAD, BN);
  p.stats = stats;
  constexpr int smem = (DPAD / 64) * BM * 128 + ST * 2 * (DPAD / 64) * BN * 128 + 1024;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    MV_CUDA(cudaFuncSetAttribute(attn_tc_kernel<DPAD, BN, ST, OCC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  }
  dim3 grid(ceil_div(a.seq_q, BM), batches * heads);
  launch_pdl(attn_tc_kernel<DPAD, BN, ST, OCC>, grid, dim3(192), smem, s, p);
}

// 40-wide heads padded to 64 (the 32x32 level): Q staged in TMEM, row sum from V's ones column
void launch64q(cudaStream_t s, const AttnSrc& a, bf16* out, float2* stats, int batches, int heads, int d) {
  constexpr int ST = 2;
  AttnParams p{};
  fill_params(p, a, out, batches, heads, d, 64, 128);
  p.stats = stats;
  constexpr int smem_q = ST * 2 * 128 * 128 + 1024;
  static bool configured_q[kMaxDevices] = {};
  if (first_use_on_device(configured_q)) {
    MV_CUDA(cudaFuncSetAttribute(attn64q_kernel<ST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_q));
  }
  dim3 grid_q(ceil_div(a.seq_q, BM), batches * heads);
  launch_pdl(attn64q_kernel<ST>, grid_q, dim3(192), smem_q, s, p);
}

void dispatch(cudaStream_t s, const AttnSrc& a, bf16* out, float2* stats, int batches, int heads, int d, int dpad) {
  MV_CHECK(d <= dpad && a.seq_q >= 1 && a.seq_kv >= 1 && batches >= 1, "attention_tc: bad arguments (need dpad >= d)");
  MV_CHECK((reinterpret_cast<uintptr_t>(a.q) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.kv) & 15) == 0 &&
               (reinterpret_cast<uintptr_t>(out) & 15) == 0 && a.ld_q % 8 == 0 && a.ld_kv % 8 == 0,
           "attention_tc: pointers must be 16-byte aligned, row pitches multiples of 8");
  // attn64q takes the row sum from V's ones column, which needs a pad column (and the K block in front of the V block);
  // d == 64 (Variant B's Transformer2D heads) runs the generic kernel, which sums the row in registers
  if (dpad == 64 && d < dpad && a.v_col0 > a.k_col0) launch64q(s, a, out, stats, batches, heads, d);
  else if (dpad == 64) launch<64, 128, 2, 2>(s, a, out, stats, batches, heads, d);   // 80 KB smem, 256 TMEM cols: 2 CTAs/SM
  else if (dpad == 128) launch<128, 128, 2, 1>(s, a, out, stats, batches, heads, d);
  else if (dpad == 192) launch<192, 64, 2, 1>(s, a, out, stats, batches, heads, d);
  else MV_CHECK(false, "attention_tc: head_dim_pad must be 64, 128 or 192");
}

// out[row, head, :] = sum_i w_i * part_i / sum_i w_i,  w_i = l_i * 2^(m_i - max_j m_j): the softmax over the union of the
// key ranges the partial launches covered.  Parts are combined in argument order (fixed: bit-stable).
__global__ void attention_merge_kernel(const bf16* __restrict__ p0, const float2* __restrict__ s0, const bf16* __restrict__ p1,
                                       const float2* __restrict__ s1, const bf16* __restrict__ p2,
                                       const float2* __restrict__ s2, int nparts, int64_t rows, int heads, int dpad,
                                       bf16* __restrict__ out) {
  const int per_row = heads * (dpad / 8);
  const int64_t total = rows * per_row;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i / per_row;
    const int head = (int)(i % per_row) / (dpad / 8);
    const bf16* parts[3] = {p0, p1, p2};
    const float2* stats[3] = {s0, s1, s2};
    float2 st[3];
    float m = -INFINITY;
    for (int k = 0; k < nparts; ++k) {
      st[k] = stats[k][row * heads + head];
      m = fmaxf(m, st[k].x);
    }
    float acc[8] = {};
    float wsum = 0.f;
    for (int k = 0; k < nparts; ++k) {
      const float w = st[k].y * exp2f(st[k].x - m);
      wsum += w;
      const uint4 u = reinterpret_cast<const uint4*>(parts[k])[i];
      const uint32_t wd[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&wd[e]));
        acc[2 * e] = fmaf(w, f.x, acc[2 * e]);
        acc[2 * e + 1] = fmaf(w, f.y, acc[2 * e + 1]);
      }
    }
    const float inv = 1.f / wsum;
    reinterpret_cast<uint4*>(out)[i] = make_uint4(pack_bf16(acc[0] * inv, acc[1] * inv), pack_bf16(acc[2] * inv, acc[3] * inv),
                                                  pack_bf16(acc[4] * inv, acc[5] * inv), pack_bf16(acc[6] * inv, acc[7] * inv));
  }
}

}  // namespace

void attention_trace_read(long long* host, int n) {
  MV_CUDA(cudaMemcpyFromSymbol(host, g_attn_trace, sizeof(long long) * n));
}

void attention_tc(cudaStream_t s, const bf16* qkv, bf16* out, int batches, int seq, int heads, int d, int dpad) {
  const int ld = 3 * heads * dpad;
  dispatch(s, AttnSrc{qkv, ld, 0, qkv, ld, heads * dpad, 2 * heads * dpad, seq, seq}, out, nullptr, batches, heads, d, dpad);
}

void attention_tc_kv(cudaStream_t s, const bf16* q, int ld_q, int q_col0, const bf16* kv, int ld_kv, int k_col0, int v_col0,
                     bf16* out, int batches, int seq_q, int seq_kv, int heads, int d, int dpad, float* stats) {
  dispatch(s, AttnSrc{q, ld_q, q_col0, kv, ld_kv, k_col0, v_col0, seq_q, seq_kv}, out, reinterpret_cast<float2*>(stats), batches,
           heads, d, dpad);
}

void attention_merge(cudaStream_t s, int nparts, const bf16* const* parts, const float* const* stats, int64_t rows, int heads,
                     int dpad, bf16* out) {
  MV_CHECK(nparts >= 1 && nparts <= 3 && dpad % 8 == 0, "attention_merge: 1..3 parts");
  const int64_t total = rows * heads * (dpad / 8);
  const int blocks = (int)std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 8);
  attention_merge_kernel<<<blocks, 256, 0, s>>>(parts[0], reinterpret_cast<const float2*>(stats[0]), nparts > 1 ? parts[1] : nullptr,
                                                nparts > 1 ? reinterpret_cast<const float2*>(stats[1]) : nullptr,
                                                nparts > 2 ? parts[2] : nullptr,
                                                nparts > 2 ? reinterpret_cast<const float2*>(stats[2]) : nullptr, nparts, rows,
                                                heads, dpad, out);
  MV_LAUNCHED();
}

}  // namespace mvldm
