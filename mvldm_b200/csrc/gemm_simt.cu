// CUDA-core implicit-GEMM (conv3x3 / 1x1 / linear) with the same descriptor and epilogues as the
// tcgen05 kernel in gemm_tc.cu.  It exists as an ON-DEVICE cross-check for the tests (select with
// MVLDM_IMPL_SIMT); the product path is gemm_tc.  fp32 accumulate over bf16 operands, fixed summation
// order (k ascending), so results are bit-stable run to run.
#include "common.cuh"

namespace mvldm {
namespace {

constexpr int BM = 64, BN = 64, BK = 16;

struct SimtParams {
  mvldm_gemm_desc d;
  int M;
};

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(256) gemm_simt_kernel(const __grid_constant__ SimtParams p) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  __shared__ float Cs[BM][BN + 1];
  const mvldm_gemm_desc& d = p.d;
  const int tid = threadIdx.x;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int tx = tid % 16, ty = tid / 16;
  const int hw = d.oh * d.ow;

  // loader assignment: row = tid / 4 (0..63), 4 consecutive k at (tid % 4) * 4
  const int lrow = tid / 4, lk = (tid % 4) * 4;
  const int am = m0 + lrow;
  const bool a_ok = am < p.M;
  int img = 0, oy = 0, ox = 0;
  if (a_ok) {
    img = am / hw;
    const int r = am - img * hw;
    oy = r / d.ow;
    ox = r - oy * d.ow;
  }
  const int bn = n0 + lrow;
  const bool b_ok = bn < d.n;

  float acc[4][4] = {};
  int seg = 0, tap = 0, c0 = 0;  // decode of the current k-block (advanced incrementally)
  for (int k0 = 0; k0 < d.k; k0 += BK) {
    // ---- A tile (implicit im2col gather) ----
    {
      const mvldm_aseg& s = d.seg[seg];
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (a_ok) {
        const int iy = oy * s.stride + s.dh[tap], ix = ox * s.stride + s.dw[tap];
        if (iy >= 0 && iy < s.sh && ix >= 0 && ix < s.sw) {
          const bf16* src = reinterpret_cast<const bf16*>(s.ptr) +
                            (((int64_t)img * s.sh + iy) * s.sw + ix) * s.ctot + s.coff[tap] + c0 + lk;
#pragma unroll
          for (int j = 0; j < 4; ++j) v[j] = __bfloat162float(src[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) As[lk + j][lrow] = v[j];
      // advance decode
      c0 += BK;
      if (c0 >= s.c) {
        c0 = 0;
        if (++tap >= s.ntaps) {
          tap = 0;
          ++seg;
        }
      }
    }
    // ---- B tile ----
    {
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (b_ok) {
        const bf16* src = reinterpret_cast<const bf16*>(d.w) + (int64_t)bn * d.k + k0 + lk;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = __bfloat162float(src[j]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) Bs[lk + j][lrow] = v[j];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue through smem so every mode gets simple, coalesced indexing ----
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
  __syncthreads();
  for (int e = tid; e < BM * BN; e += 256) {
    const int r = e / BN, c = e % BN;
    const int m = m0 + r, n = n0 + c;
    if (m >= p.M || n >= d.n) continue;
    const int im = m / hw;
    auto val = [&](int cc) -> float {
      const int nn = n0 + cc;
      float v = Cs[r][cc];
      if (d.bias) v += d.bias[nn];
      if (d.rowvec) v += d.rowvec[(int64_t)im * d.rowvec_ld + nn];
      return v;
    };
    if (d.mode == 0 || d.mode == 3 || d.mode == 5) {
      float v = val(c);
      if (d.residual) v += __bfloat162float(reinterpret_cast<const bf16*>(d.residual)[(int64_t)m * d.res_ld + n]);
      if (d.mode == 3) v = v / (1.f + expf(-v));
      if (d.mode == 5) v = gelu_exact(v);
      reinterpret_cast<bf16*>(d.out)[(int64_t)m * d.ldo + n] = __float2bfloat16(v);
    } else if (d.mode == 4) {
      reinterpret_cast<float*>(d.out)[(int64_t)m * d.ldo + n] = val(c);
    } else if (d.mode == 1) {
      if ((c & 15) < 8) {  // columns [16j,16j+8) are x, [16j+8,16j+16) the matching gates
        const float x = val(c), g = val(c + 8);
        const int oc = (n / 16) * 8 + (n & 7);
        reinterpret_cast<bf16*>(d.out)[(int64_t)m * d.ldo + oc] = __float2bfloat16(x * gelu_exact(g));
      }
    } else {
      if (n < d.n_valid) {
        const int pix = m - im * hw;
        reinterpret_cast<float*>(d.out)[((int64_t)im * d.n_valid + n) * hw + pix] = val(c);
      }
    }
  }
}

}  // namespace

void gemm_simt(cudaStream_t s, const mvldm_gemm_desc& d) {
  int ktot = 0;
  for (int i = 0; i < d.nseg; ++i) {
    MV_CHECK(d.seg[i].c % BK == 0, "gemm_simt: segment channels must be a multiple of 16");
    ktot += d.seg[i].c * d.seg[i].ntaps;
  }
  MV_CHECK(ktot == d.k, "gemm_simt: K mismatch between segments and weights");
  MV_CHECK(d.mode != 1 || d.n % 64 == 0, "gemm_simt: GEGLU needs N % 64 == 0");
  SimtParams p;
  p.d = d;
  p.M = d.n_img * d.oh * d.ow;
  dim3 grid(ceil_div(p.M, BM), ceil_div(d.n, BN));
  gemm_simt_kernel<<<grid, 256, 0, s>>>(p);
  MV_LAUNCHED();
}

}  // namespace mvldm
