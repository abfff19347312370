// CUDA-core streaming-softmax self-attention; the on-device cross-check for attn_tc.cu (tests only).
// Same I/O contract: packed head-padded q|k|v in, head-padded output out, fp32 scores and softmax
// (mvdream/attention.py:185-203), P and O kept in fp32 here.
#include "common.cuh"

namespace mvldm {
namespace {

constexpr int QT = 32;       // queries per CTA (8 per warp)
constexpr int KT = 64;       // keys per smem tile
constexpr int MAXD32 = 6;    // dpad <= 192

__global__ void __launch_bounds__(128) attn_simt_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int seq,
                                                        int heads, int d, int dpad, float scale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int pitch = dpad + 2;
  bf16* Ks = reinterpret_cast<bf16*>(smem_raw);
  bf16* Vs = Ks + KT * pitch;
  float* Qs = reinterpret_cast<float*>(Vs + KT * pitch);  // [QT][dpad]

  const int bh = blockIdx.y, batch = bh / heads, head = bh % heads;
  const int q0 = blockIdx.x * QT;
  const int ld = 3 * heads * dpad;
  const bf16* base = qkv + (int64_t)batch * seq * ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nd = dpad / 32;

  for (int i = tid; i < QT * dpad; i += 128) {
    const int q = i / dpad, j = i % dpad;
    Qs[i] = (q0 + q < seq) ? __bfloat162float(base[(int64_t)(q0 + q) * ld + head * dpad + j]) : 0.f;
  }
  float m_run[8], l_run[8], o[8][MAXD32];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    m_run[q] = -INFINITY;
    l_run[q] = 0.f;
#pragma unroll
    for (int i = 0; i < MAXD32; ++i) o[q][i] = 0.f;
  }
  for (int k0 = 0; k0 < seq; k0 += KT) {
    __syncthreads();
    for (int i = tid; i < KT * dpad; i += 128) {
      const int r = i / dpad, j = i % dpad;
      bf16 kv = __float2bfloat16(0.f), vv = kv;
      if (k0 + r < seq) {
        kv = base[(int64_t)(k0 + r) * ld + (heads + head) * dpad + j];
        vv = base[(int64_t)(k0 + r) * ld + (2 * heads + head) * dpad + j];
      }
      Ks[r * pitch + j] = kv;
      Vs[r * pitch + j] = vv;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float* qv = Qs + (warp * 8 + q) * dpad;
      float s0 = 0.f, s1 = 0.f;
      const __nv_bfloat162* kr0 = reinterpret_cast<const __nv_bfloat162*>(Ks + lane * pitch);
      const __nv_bfloat162* kr1 = reinterpret_cast<const __nv_bfloat162*>(Ks + (lane + 32) * pitch);
      for (int j = 0; j < d / 2; ++j) {
        const float2 a = __bfloat1622float2(kr0[j]), b = __bfloat1622float2(kr1[j]);
        const float qa = qv[2 * j], qb = qv[2 * j + 1];
        s0 = fmaf(qa, a.x, fmaf(qb, a.y, s0));
        s1 = fmaf(qa, b.x, fmaf(qb, b.y, s1));
      }
      s0 = (k0 + lane < seq) ? s0 * scale : -INFINITY;
      s1 = (k0 + lane + 32 < seq) ? s1 * scale : -INFINITY;
      float tm = fmaxf(s0, s1);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, off));
      const float m_new = fmaxf(m_run[q], tm);
      const float corr = __expf(m_run[q] - m_new);
      const float p0 = __expf(s0 - m_new), p1 = __expf(s1 - m_new);
      float ps = p0 + p1;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l_run[q] = l_run[q] * corr + ps;
      m_run[q] = m_new;
#pragma unroll
      for (int i = 0; i < MAXD32; ++i) o[q][i] *= corr;
      for (int kk = 0; kk < 32; ++kk) {
        const float pa = __shfl_sync(0xffffffffu, p0, kk), pb = __shfl_sync(0xffffffffu, p1, kk);
#pragma unroll
        for (int i = 0; i < MAXD32; ++i) {
          if (i < nd) {
            o[q][i] = fmaf(pa, __bfloat162float(Vs[kk * pitch + lane + 32 * i]), o[q][i]);
            o[q][i] = fmaf(pb, __bfloat162float(Vs[(kk + 32) * pitch + lane + 32 * i]), o[q][i]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    const int qi = q0 + warp * 8 + q;
    if (qi >= seq) continue;
    const float inv = 1.f / l_run[q];
    bf16* dst = out + ((int64_t)batch * seq + qi) * (heads * dpad) + head * dpad;
#pragma unroll
    for (int i = 0; i < MAXD32; ++i) {
      if (i < nd) {
        const int j = lane + 32 * i;
        dst[j] = __float2bfloat16(j < d ? o[q][i] * inv : 0.f);
      }
    }
  }
}

}  // namespace

void attention_simt(cudaStream_t s, const bf16* qkv, bf16* out, int batches, int seq, int heads, int d, int dpad) {
  MV_CHECK(dpad % 32 == 0 && dpad <= 32 * MAXD32 && d <= dpad && d % 2 == 0, "attention_simt: unsupported head dim");
  const size_t smem = (size_t)2 * KT * (dpad + 2) * sizeof(bf16) + (size_t)QT * dpad * sizeof(float);
  MV_CUDA(cudaFuncSetAttribute(attn_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(ceil_div(seq, QT), batches * heads);
  attn_simt_kernel<<<grid, 128, smem, s>>>(qkv, out, seq, heads, d, dpad, 1.f / sqrtf((float)d));
  MV_LAUNCHED();
}

}  // namespace mvldm
