// GroupNorm(+SiLU), LayerNorm, nearest upsample, input im2col and timestep sinusoid.
// bf16 NHWC activations ([image, h, w, channel]), fp32 statistics, fixed reduction orders (bit-stable).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace mvldm {

namespace {

// MUFU.EX2 + MUFU.RCP: the IEEE division this replaces was ~10x the instructions of the rest of a GroupNorm element
__device__ __forceinline__ float silu_f(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// eight bf16 moved as ONE 16-byte access (a struct of four __nv_bfloat162 is copied member by member: 4 x 32-bit)
struct alignas(16) bf16x8 {
  uint4 u;
  __device__ __forceinline__ __nv_bfloat162 h(int i) const {
    const uint32_t w = i == 0 ? u.x : (i == 1 ? u.y : (i == 2 ? u.z : u.w));
    return *reinterpret_cast<const __nv_bfloat162*>(&w);
  }
};
__device__ __forceinline__ uint32_t bf162_bits(float a, float b) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&t);
}

__device__ __forceinline__ void load8(const bf16* p, float* f) {
  bf16x8 r = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(r.h(i));
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
  bf16x8 r;
  r.u = make_uint4(bf162_bits(f[0], f[1]), bf162_bits(f[2], f[3]), bf162_bits(f[4], f[5]), bf162_bits(f[6], f[7]));
  *reinterpret_cast<bf16x8*>(p) = r;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deterministic block sum (fixed shuffle tree, fixed warp order); result broadcast to all threads
__device__ __forceinline__ float block_sum(float v, float* red /* >= 33 floats */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---------------------------------------------------------------------------------------------
// conv_in operand: explicit im2col of the fp32 NCHW denoiser input (K1 + K3 of SURVEY.md §2.2).
// out[m, tap*cin + c] = latents[img, c, y+r-1, x+s-1] (zero outside / for k >= 9*cin), m = (img, y, x)
// ---------------------------------------------------------------------------------------------
// `premix` (optional, fp32 [cin*cin + cin]): a 1x1 convolution y = W x + b applied to every pixel before the 3x3 gather (the VAE's
// post_quant_conv in front of decoder.conv_in: its bias must not leak into the zero padding, so it cannot be folded into the
// 3x3 weights)
__global__ void im2col_input_kernel(const float* __restrict__ x, int n_img, int cin, int h, int w, int kpad,
                                    const float* __restrict__ premix, bf16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int kv = kpad / 8;  // one thread = eight consecutive k of one output pixel = one 16-byte store
  const int total = n_img * h * w * kv;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
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
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          if (premix) {
            float a = premix[cin * cin + c];
            for (int e = 0; e < cin; ++e) a = fmaf(premix[c * cin + e], x[((img * cin + e) * h + yy) * w + xx], a);
            v[j] = a;
          } else {
            v[j] = x[((img * cin + c) * h + yy) * w + xx];
          }
        }
      }
    }
    store8(out + (int64_t)i * 8, v);
  }
}

// out[r, :] = softmax(scale * in[r, :]) for `rows` rows of `cols` fp32 scores -> bf16 (one warp per row, the row in registers;
// fp32 max / sum like diffusers' upcast_softmax): the VAE's single-head 512-wide mid-block attention runs as plain GEMMs
template <int PER_LANE>
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ in, int64_t rows, int cols, float scale,
                                                           bf16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* src = in + row * cols;
  float v[PER_LANE];
  float mx = -INFINITY;
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) {
    const int c = j * 32 + lane;
    v[j] = c < cols ? src[c] * scale : -INFINITY;
    mx = fmaxf(mx, v[j]);
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) {
    v[j] = __expf(v[j] - mx);
    sum += v[j];
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.f / sum;
  bf16* dst = out + row * cols;
#pragma unroll
  for (int j = 0; j < PER_LANE; ++j) {
    const int c = j * 32 + lane;
    if (c < cols) dst[c] = __float2bfloat16(v[j] * inv);
  }
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0): [cos | sin], fp32
template <class OutT>
__global__ void sinusoid_kernel(const int64_t* __restrict__ t, int n, int dim, OutT* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int half = dim / 2;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * half) return;
  const int r = i / half, j = i - r * half;
  const float freq = expf(-logf(10000.f) * (float)j / (float)half);
  const float arg = (float)t[r] * freq;
  out[(int64_t)r * dim + j] = (OutT)cosf(arg);
  out[(int64_t)r * dim + half + j] = (OutT)sinf(arg);
}

// ---------------------------------------------------------------------------------------------
// GroupNorm over NHWC bf16.  Pass 1 (gn_partial): grid (pixel chunks, images); every thread owns one
// 8-channel vector column and walks the chunk's pixels with 16-byte loads, per-channel (sum, sum of squares)
// are combined through shared memory in a fixed order and folded to per-group partials
// [image][chunk][group][2] (no atomics: bit-stable).  Pass 2 (gn_apply): the first warp folds the partials of
// its image to (mean, rstd) in double, then the CTA normalises * gamma + beta (+SiLU) with 16-byte vectors.
// The two sources implement GroupNorm over torch.cat((hidden, skip), dim=1) without materialising the cat.
// ---------------------------------------------------------------------------------------------
constexpr int GN_MAXC = 2560;
constexpr int GN_MAXP = 64;

__global__ void __launch_bounds__(256) gn_partial_kernel(const bf16* __restrict__ x0, int c0, const bf16* __restrict__ x1,
                                                         int c1, int hw, int groups, int pix, float* __restrict__ partials) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float red_s[GN_MAXC], red_q[GN_MAXC];
  const int C = c0 + c1, ncv = C / 8, cg = C / groups;
  const int img = blockIdx.y, chunk = blockIdx.x, P = gridDim.x;
  const int t = threadIdx.x;
  const int lanes_p = ncv <= 256 ? 256 / ncv : 1;
  const int pl = ncv <= 256 ? t / ncv : 0;
  const int64_t pix0 = (int64_t)img * hw + (int64_t)chunk * pix;
  if (pl < lanes_p) {
    for (int cv = ncv <= 256 ? t % ncv : t; cv < ncv; cv += 256) {
      float sa[8], qa[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) sa[j] = qa[j] = 0.f;
      const int c = cv * 8;
      const bf16* src = c < c0 ? x0 + pix0 * c0 + c : x1 + pix0 * c1 + (c - c0);
      const int64_t pitch = c < c0 ? c0 : c1;
      int pp = pl;
      for (; pp + 3 * lanes_p < pix; pp += 4 * lanes_p) {  // 4 independent 16-byte loads in flight per thread
        float f[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) load8(src + (int64_t)(pp + u * lanes_p) * pitch, f[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            sa[j] += f[u][j];
            qa[j] = fmaf(f[u][j], f[u][j], qa[j]);
          }
      }
      for (; pp < pix; pp += lanes_p) {
        float f[8];
        load8(src + (int64_t)pp * pitch, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          sa[j] += f[j];
          qa[j] = fmaf(f[j], f[j], qa[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        red_s[pl * C + c + j] = sa[j];
        red_q[pl * C + c + j] = qa[j];
      }
      if (ncv <= 256) break;
    }
  }
  __syncthreads();
  if (t < groups) {
    float S = 0.f, Q = 0.f;
    for (int l = 0; l < lanes_p; ++l)
      for (int c = t * cg; c < (t + 1) * cg; ++c) {
        S += red_s[l * C + c];
        Q += red_q[l * C + c];
      }
    float* o = partials + (((int64_t)img * P + chunk) * groups + t) * 2;
    o[0] = S;
    o[1] = Q;
  }
}

__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x0, int c0, const bf16* __restrict__ x1,
                                                       int c1, int hw, int groups, float eps, int P,
                                                       const float* __restrict__ partials, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int silu, bf16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ float s_mean[64], s_rstd[64];
  __shared__ double s_part[4][64][2];
  const int C = c0 + c1, cg = C / groups, cv = C / 8;
  const int img = blockIdx.y;
  {  // fold the chunk partials: 4 slices of chunks in parallel (independent loads), then a fixed-order sum
    const int g = threadIdx.x & 63, slice = threadIdx.x >> 6;
    if (g < groups) {
      double S = 0.0, Q = 0.0;
#pragma unroll 4
      for (int k = slice; k < P; k += 4) {
        const float2 o = *reinterpret_cast<const float2*>(partials + (((int64_t)img * P + k) * groups + g) * 2);
        S += (double)o.x;
        Q += (double)o.y;
      }
      s_part[slice][g][0] = S;
      s_part[slice][g][1] = Q;
    }
  }
  __syncthreads();
  if (threadIdx.x < groups) {
    const int g = threadIdx.x;
    const double S = ((s_part[0][g][0] + s_part[1][g][0]) + s_part[2][g][0]) + s_part[3][g][0];
    const double Q = ((s_part[0][g][1] + s_part[1][g][1]) + s_part[2][g][1]) + s_part[3][g][1];
    const double cnt = (double)hw * cg;
    const double mean = S / cnt;
    double var = Q / cnt - mean * mean;
    var = var < 0.0 ? 0.0 : var;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  const int64_t per_img = (int64_t)hw * cv;
  const int64_t lo = per_img * blockIdx.x / gridDim.x, hi = per_img * (blockIdx.x + 1) / gridDim.x;
#pragma unroll 2
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const int c = (int)(i % cv) * 8;
    const int64_t pixel = (int64_t)img * hw + i / cv;
    float f[8];
    if (c < c0) load8(x0 + pixel * c0 + c, f);
    else load8(x1 + pixel * c1 + (c - c0), f);
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (c + j) / cg;
      const float v = (f[j] - s_mean[g]) * s_rstd[g] * gg[j] + bb[j];
      f[j] = silu ? silu_f(v) : v;
    }
    store8(out + pixel * C + c, f);
  }
}

// ---------------------------------------------------------------------------------------------
// Single-launch GroupNorm: groups are independent, so one CTA owns (image, block of whole groups): it streams
// its [hw x cb] channel slab once (kept in shared memory when it fits), reduces per-channel partials across
// pixel lanes and then per group in a fixed order (bit-stable), and normalises the slab.  cb = lcm(group
// width, 8) channels so that every access is a 16-byte vector.
// ---------------------------------------------------------------------------------------------
constexpr int GNB_THREADS = 512;

__global__ void __launch_bounds__(GNB_THREADS) gn_block_kernel(const bf16* __restrict__ x0, int c0,
                                                               const bf16* __restrict__ x1, int c1, int hw, int groups,
                                                               float eps, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, int silu, int cb,
                                                               int cache, bf16* __restrict__ out) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();  // CS CTAs along x share one (image, channel block): pixel split
  extern __shared__ __align__(16) uint8_t gnb_smem[];
  __shared__ float red_s[GNB_THREADS * 8], red_q[GNB_THREADS * 8];
  __shared__ float chan_s[GNB_THREADS], chan_q[GNB_THREADS];
  __shared__ float my_part[16][2];
  __shared__ float s_mean[16], s_rstd[16];
  pdl_wait();
  pdl_launch_dependents();
  bf16* slab = reinterpret_cast<bf16*>(gnb_smem);
  const int CS = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int hw_all = hw;
  hw = hw_all / CS;                                // pixels this CTA owns
  const int C = c0 + c1, cgn = C / groups, nv = cb / 8, gb = cb / cgn;
  const int img = blockIdx.y, ch0 = (blockIdx.x / CS) * cb, t = threadIdx.x;
  const int lanes_p = min(GNB_THREADS / nv, hw);  // pixel lanes (no more than there are pixels)
  const int cv = t % nv, pl = t / nv;
  const int c = ch0 + cv * 8;
  const int64_t pix0 = (int64_t)img * hw_all + (int64_t)crank * hw;
  const bf16* src = c < c0 ? x0 + pix0 * c0 + c : x1 + pix0 * c1 + (c - c0);
  const int64_t pitch = c < c0 ? c0 : c1;
  if (pl < lanes_p) {
    float sa[8], qa[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sa[j] = qa[j] = 0.f;
    int pp = pl;
    for (; pp + 3 * lanes_p < hw; pp += 4 * lanes_p) {
      bf16x8 raw[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) raw[u] = *reinterpret_cast<const bf16x8*>(src + (int64_t)(pp + u * lanes_p) * pitch);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (cache) *reinterpret_cast<bf16x8*>(slab + (int64_t)(pp + u * lanes_p) * cb + cv * 8) = raw[u];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __bfloat1622float2(raw[u].h(j));
          sa[2 * j] += f.x; sa[2 * j + 1] += f.y;
          qa[2 * j] = fmaf(f.x, f.x, qa[2 * j]); qa[2 * j + 1] = fmaf(f.y, f.y, qa[2 * j + 1]);
        }
      }
    }
    for (; pp < hw; pp += lanes_p) {
      const bf16x8 raw = *reinterpret_cast<const bf16x8*>(src + (int64_t)pp * pitch);
      if (cache) *reinterpret_cast<bf16x8*>(slab + (int64_t)pp * cb + cv * 8) = raw;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(raw.h(j));
        sa[2 * j] += f.x; sa[2 * j + 1] += f.y;
        qa[2 * j] = fmaf(f.x, f.x, qa[2 * j]); qa[2 * j + 1] = fmaf(f.y, f.y, qa[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {  // [pixel lane][channel in block]
      red_s[pl * cb + cv * 8 + j] = sa[j];
      red_q[pl * cb + cv * 8 + j] = qa[j];
    }
  }
  __syncthreads();
  {  // fixed-order fold: (1) per channel over the pixel lanes, `parts` threads per channel; (2) per group, one warp
    const int parts = GNB_THREADS / cb;
    const int ch = t % cb, part = t / cb;
    if (part < parts) {
      float S = 0.f, Q = 0.f;
      for (int l = part; l < lanes_p; l += parts) {
        S += red_s[l * cb + ch];
        Q += red_q[l * cb + ch];
      }
      chan_s[part * cb + ch] = S;
      chan_q[part * cb + ch] = Q;
    }
    __syncthreads();
    const int warp = t >> 5, lane = t & 31;
    if (warp < gb) {
      float S = 0.f, Q = 0.f;
      const int n = parts * cgn;
      for (int i = lane; i < n; i += 32) {
        const int pt = i / cgn, cc = warp * cgn + i % cgn;
        S += chan_s[pt * cb + cc];
        Q += chan_q[pt * cb + cc];
      }
      S = warp_sum(S);
      Q = warp_sum(Q);
      if (lane == 0) {
        my_part[warp][0] = S;
        my_part[warp][1] = Q;
      }
    }
  }
  if (CS > 1) cluster.sync(); else __syncthreads();
  if (t < gb) {  // fold the cluster's partials in rank order (bit-stable) through distributed shared memory
    float S = 0.f, Q = 0.f;
    for (int r = 0; r < CS; ++r) {
      const float* peer = CS > 1 ? cluster.map_shared_rank(&my_part[0][0], r) : &my_part[0][0];
      S += peer[2 * t];
      Q += peer[2 * t + 1];
    }
    const float cnt = (float)hw_all * (float)cgn;
    const float mean = S / cnt;
    const float var = fmaxf(Q / cnt - mean * mean, 0.f);
    s_mean[t] = mean;
    s_rstd[t] = rsqrtf(var + eps);
  }
  if (CS > 1) cluster.sync(); else __syncthreads();  // also keeps my_part alive until every peer has read it
  if (pl < lanes_p) {
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int g = (cv * 8 + j) / cgn;
      sc[j] = s_rstd[g] * gg[j];
      sh[j] = bb[j] - s_mean[g] * sc[j];
    }
    bf16* dst = out + pix0 * C + c;
#pragma unroll 4
    for (int pp = pl; pp < hw; pp += lanes_p) {
      float f[8];
      if (cache) load8(slab + (int64_t)pp * cb + cv * 8, f);
      else load8(src + (int64_t)pp * pitch, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = fmaf(f[j], sc[j], sh[j]);
        f[j] = silu ? silu_f(v) : v;
      }
      store8(dst + (int64_t)pp * C, f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm, register-resident: one (image, block of CB = 8*NV channels = whole groups) item is owned by `tpi`
// threads, thread r holding all CB channels of pixels r, r + tpi, ... (P of them) in registers.  Every load is in flight
// before the first use, the pixel reduction is a shuffle tree plus (tpi > 32) ONE __syncthreads, and the normalised
// values are written straight from the registers: no staging slab, no second read, no cluster barrier.  Small images
// pack several items into a CTA (tpi = 16 for the 4x4 level).  Fixed reduction order -> bit-stable.
// ---------------------------------------------------------------------------------------------
template <int NV, int P, int CGN>
__global__ void __launch_bounds__(512) gn_reg_kernel(const bf16* __restrict__ x0, int c0, const bf16* __restrict__ x1,
                                                     int c1, int hw, float eps, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, int silu, int tpi, int n_items,
                                                     bf16* __restrict__ out) {
  constexpr int CB = NV * 8, GB = CB / CGN;
  static_assert(CB % CGN == 0, "channel block must hold whole groups");
  __shared__ float red[16][2 * GB];
  pdl_wait();
  pdl_launch_dependents();
  const int C = c0 + c1, nblk = C / CB;
  const int t = threadIdx.x, ipc = blockDim.x / tpi;
  const int item = blockIdx.x * ipc + t / tpi, r = t % tpi;
  const bool live = item < n_items;
  const int img = live ? item / nblk : 0, ch0 = live ? (item % nblk) * CB : 0;
  bf16x8 raw[P][NV];
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const int64_t pix = (int64_t)img * hw + r + j * tpi;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = ch0 + v * 8;
      const bf16* src = c < c0 ? x0 + pix * c0 + c : x1 + pix * c1 + (c - c0);
      if (live) raw[j][v] = *reinterpret_cast<const bf16x8*>(src);
      else raw[j][v].u = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float S[GB], Q[GB], K[GB];   // sums about a per-group pivot (the group's first channel at the image's first pixel), see gn_flat
#pragma unroll
  for (int g = 0; g < GB; ++g) {
    S[g] = Q[g] = 0.f;
    const int c = ch0 + g * CGN;
    const int64_t pix0 = (int64_t)img * hw;
    K[g] = __bfloat162float(c < c0 ? x0[pix0 * c0 + c] : x1[pix0 * c1 + (c - c0)]);
  }
#pragma unroll
  for (int j = 0; j < P; ++j)
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(raw[j][v].h(e));
        const int ga = (v * 8 + 2 * e) / CGN, gbb = (v * 8 + 2 * e + 1) / CGN;  // compile-time after unrolling
        const float dx = f.x - K[ga], dy = f.y - K[gbb];
        S[ga] += dx; Q[ga] = fmaf(dx, dx, Q[ga]);
        S[gbb] += dy; Q[gbb] = fmaf(dy, dy, Q[gbb]);
      }
  const int span = tpi < 32 ? tpi : 32;
#pragma unroll
  for (int g = 0; g < GB; ++g)
    for (int o = span >> 1; o > 0; o >>= 1) {
      S[g] += __shfl_xor_sync(0xffffffffu, S[g], o);
      Q[g] += __shfl_xor_sync(0xffffffffu, Q[g], o);
    }
  if (tpi > 32) {
    const int warp = t >> 5, lane = t & 31;
    if (lane == 0) {
#pragma unroll
      for (int g = 0; g < GB; ++g) {
        red[warp][2 * g] = S[g];
        red[warp][2 * g + 1] = Q[g];
      }
    }
    __syncthreads();
    const int w0 = (t / tpi) * (tpi >> 5), nw = tpi >> 5;
#pragma unroll
    for (int g = 0; g < GB; ++g) {
      float s = 0.f, q = 0.f;
      for (int w = 0; w < nw; ++w) {
        s += red[w0 + w][2 * g];
        q += red[w0 + w][2 * g + 1];
      }
      S[g] = s;
      Q[g] = q;
    }
  }
  if (!live) return;
  const float inv_cnt = 1.f / ((float)hw * (float)CGN);
  float mean[GB], rstd[GB];
#pragma unroll
  for (int g = 0; g < GB; ++g) {
    const float dm = S[g] * inv_cnt;   // mean - pivot
    mean[g] = K[g] + dm;
    rstd[g] = rsqrtf(fmaxf(Q[g] * inv_cnt - dm * dm, 0.f) + eps);
  }
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    const int c = ch0 + v * 8;
    const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
    const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float sc[8], sh[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int g = (v * 8 + e) / CGN;
      sc[e] = rstd[g] * gg[e];
      sh[e] = bb[e] - mean[g] * sc[e];
    }
#pragma unroll
    for (int j = 0; j < P; ++j) {
      float f[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 x = __bfloat1622float2(raw[j][v].h(e));
        f[2 * e] = x.x;
        f[2 * e + 1] = x.y;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float y = fmaf(f[e], sc[e], sh[e]);
        f[e] = silu ? silu_f(y) : y;
      }
      store8(out + ((int64_t)img * hw + r + j * tpi) * C + c, f);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// GroupNorm, register-resident, coalesced: one CTA per (image, 40-channel block = whole groups).  The block's
// hw x 5 sixteen-byte vectors are dealt round-robin to T threads (T a multiple of 5, so a thread always sees the same
// 8 channels and consecutive threads read consecutive vectors), R = hw*5/T vectors per thread, all in flight at once.
// Per-channel sums stay in registers; each thread folds its 8 channels into the (at most two) groups they belong to,
// a shuffle tree + one __syncthreads sums over the CTA, and the output is written from the registers.  ~3x fewer
// instructions per element than the slab kernel, which is what bounds GroupNorm at large batch.
// ---------------------------------------------------------------------------------------------
template <int NV, int R, int CGN>
__global__ void __launch_bounds__(NV == 5 ? 320 : 480, NV == 5 ? 3 : 2)
    gn_flat_kernel(const bf16* __restrict__ x0, int c0, const bf16* __restrict__ x1, int c1, int hw, float eps,
                   const float* __restrict__ gamma, const float* __restrict__ beta, int silu, bf16* __restrict__ out) {
  namespace cg = cooperative_groups;
  constexpr int CB = NV * 8, GB = CB / CGN;   // NV = 5: 40-channel blocks (groups of 10/20/40); NV = 15: 120 (30/60)
  static_assert(CB % CGN == 0 && CGN >= 8, "a block holds whole groups; a vector spans at most two");
  __shared__ float red[16][2 * GB];
  __shared__ float part[2 * GB];   // this CTA's sums; peers of the cluster read it through DSMEM
  __shared__ float stat[2 * GB];
  pdl_wait();
  pdl_launch_dependents();
  cg::cluster_group cluster = cg::this_cluster();   // CS CTAs along x split the pixels of one (image, channel block)
  const int CS = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int C = c0 + c1, nblk = C / CB, T = blockDim.x;
  const int t = threadIdx.x, item = blockIdx.x / CS, img = item / nblk, ch0 = (item % nblk) * CB;
  const int cv = t % NV, pstep = T / NV;   // vector j of this thread: pixel p0 + j * pstep, channels c..c+7
  const int p0 = crank * (hw / CS) + t / NV;
  const int c = ch0 + cv * 8;
  const bool first = c < c0;
  const bf16* src = first ? x0 + ((int64_t)img * hw + p0) * c0 + c : x1 + ((int64_t)img * hw + p0) * c1 + (c - c0);
  const int64_t sstep = (int64_t)pstep * (first ? c0 : c1);
  bf16x8 raw[R];
#pragma unroll
  for (int j = 0; j < R; ++j) raw[j] = *reinterpret_cast<const bf16x8*>(src + j * sstep);
  // channels cv*8 .. cv*8+7 belong to at most two groups (CGN >= 8): g_lo for e < eb, g_lo + 1 from eb on
  const int g_lo = (cv * 8) / CGN, eb = (g_lo + 1) * CGN - cv * 8;
  // Sums are taken about a pivot K per group - the group's first channel at the image's first pixel, the same value in
  // every CTA of the cluster: mean = K + S1/n, var = S2/n - (S1/n)^2 then cancels against (mean - K)^2 = O(var) instead
  // of mean^2, so |mean| >> std (real checkpoints) costs no precision.
  float k_lo, k_hi;
  {
    const int c_lo = ch0 + g_lo * CGN, c_hi = ch0 + min(g_lo + 1, GB - 1) * CGN;
    const int64_t pix0 = (int64_t)img * hw;
    k_lo = __bfloat162float(c_lo < c0 ? x0[pix0 * c0 + c_lo] : x1[pix0 * c1 + (c_lo - c0)]);
    k_hi = __bfloat162float(c_hi < c0 ? x0[pix0 * c0 + c_hi] : x1[pix0 * c1 + (c_hi - c0)]);
  }
  float sa[8], qa[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) sa[e] = qa[e] = 0.f;
#pragma unroll
  for (int j = 0; j < R; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __bfloat1622float2(raw[j].h(e));
      const float dx = f.x - (2 * e < eb ? k_lo : k_hi), dy = f.y - (2 * e + 1 < eb ? k_lo : k_hi);
      sa[2 * e] += dx; qa[2 * e] = fmaf(dx, dx, qa[2 * e]);
      sa[2 * e + 1] += dy; qa[2 * e + 1] = fmaf(dy, dy, qa[2 * e + 1]);
    }
  float s_lo = 0.f, q_lo = 0.f, s_hi = 0.f, q_hi = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool lo = e < eb;
    s_lo += lo ? sa[e] : 0.f; q_lo += lo ? qa[e] : 0.f;
    s_hi += lo ? 0.f : sa[e]; q_hi += lo ? 0.f : qa[e];
  }
  float S[GB], Q[GB];
#pragma unroll
  for (int g = 0; g < GB; ++g) {
    S[g] = (g == g_lo ? s_lo : 0.f) + (g == g_lo + 1 ? s_hi : 0.f);
    Q[g] = (g == g_lo ? q_lo : 0.f) + (g == g_lo + 1 ? q_hi : 0.f);
  }
#pragma unroll
  for (int g = 0; g < GB; ++g)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      S[g] += __shfl_xor_sync(0xffffffffu, S[g], o);
      Q[g] += __shfl_xor_sync(0xffffffffu, Q[g], o);
    }
  const int warp = t >> 5, lane = t & 31, nw = T >> 5;
  if (lane == 0) {
#pragma unroll
    for (int g = 0; g < GB; ++g) {
      red[warp][2 * g] = S[g];
      red[warp][2 * g + 1] = Q[g];
    }
  }
  __syncthreads();
  if (t < 2 * GB) {
    float a = 0.f;
    for (int w = 0; w < nw; ++w) a += red[w][t];
    part[t] = a;
    if (CS == 1) stat[t] = a;
  }
  if (CS > 1) {
    cluster.sync();
    if (t < 2 * GB) {  // rank order: bit-stable
      float a = 0.f;
      for (int r = 0; r < CS; ++r) a += cluster.map_shared_rank(&part[0], r)[t];
      stat[t] = a;
    }
    cluster.sync();    // also keeps `part` alive until every peer has read it
  } else {
    __syncthreads();
  }
  const float inv_cnt = 1.f / ((float)hw * (float)CGN);
  const int g_hi = min(g_lo + 1, GB - 1);
  const float d_lo = stat[2 * g_lo] * inv_cnt, d_hi = stat[2 * g_hi] * inv_cnt;   // mean - pivot
  const float mean_lo = k_lo + d_lo, mean_hi = k_hi + d_hi;
  const float rstd_lo = rsqrtf(fmaxf(stat[2 * g_lo + 1] * inv_cnt - d_lo * d_lo, 0.f) + eps);
  const float rstd_hi = rsqrtf(fmaxf(stat[2 * g_hi + 1] * inv_cnt - d_hi * d_hi, 0.f) + eps);
  const float4 g0 = *reinterpret_cast<const float4*>(gamma + c), g1 = *reinterpret_cast<const float4*>(gamma + c + 4);
  const float4 b0 = *reinterpret_cast<const float4*>(beta + c), b1 = *reinterpret_cast<const float4*>(beta + c + 4);
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const bool lo = e < eb;
    sc[e] = (lo ? rstd_lo : rstd_hi) * gg[e];
    sh[e] = bb[e] - (lo ? mean_lo : mean_hi) * sc[e];
  }
  bf16* dst = out + ((int64_t)img * hw + p0) * C + c;
  const int64_t dstep = (int64_t)pstep * C;
#pragma unroll
  for (int j = 0; j < R; ++j) {
    float f[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 x = __bfloat1622float2(raw[j].h(e));
      f[2 * e] = x.x;
      f[2 * e + 1] = x.y;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float y = fmaf(f[e], sc[e], sh[e]);
      f[e] = silu ? silu_f(y) : y;
    }
    store8(dst + j * dstep, f);
  }
}

// LayerNorm over the channel dim, one warp per token, row cached in registers (C <= 32*8*MAXV).
template <int MAXV>
__global__ void layernorm_kernel(const bf16* __restrict__ x, int rows, int c, float eps, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, bf16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nv = c / 8;
  float f[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      load8(x + (int64_t)row * c + v * 8, f[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += f[i][j];
    }
  }
  const float mean = warp_sum(s) / (float)c;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    if (lane + 32 * i < nv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = f[i][j] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)c + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int v = lane + 32 * i;
    if (v < nv) {
      const float4 g0 = *reinterpret_cast<const float4*>(gamma + v * 8), g1 = *reinterpret_cast<const float4*>(gamma + v * 8 + 4);
      const float4 b0 = *reinterpret_cast<const float4*>(beta + v * 8), b1 = *reinterpret_cast<const float4*>(beta + v * 8 + 4);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (f[i][j] - mean) * rstd * gg[j] + bb[j];
      store8(out + (int64_t)row * c + v * 8, o);
    }
  }
}

__global__ void upsample2x_kernel(const bf16* __restrict__ x, int n_img, int h, int w, int c, bf16* __restrict__ out) {
  pdl_wait();
  pdl_launch_dependents();
  const int cv = c / 8;
  const int64_t total = (int64_t)n_img * 4 * h * w * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % cv);
    int64_t p = i / cv;
    const int ox = (int)(p % (2 * w));
    p /= 2 * w;
    const int oy = (int)(p % (2 * h));
    const int img = (int)(p / (2 * h));
    const int4 val = *reinterpret_cast<const int4*>(x + (((int64_t)img * h + oy / 2) * w + ox / 2) * c + v * 8);
    *reinterpret_cast<int4*>(out + i * 8) = val;
  }
}

inline int grid_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void softmax_rows(cudaStream_t s, const float* in, int64_t rows, int cols, float scale, bf16* out) {
  MV_CHECK(cols >= 1 && cols <= 4096, "softmax_rows: 1..4096 columns");
  const dim3 grid((unsigned)ceil_div64(rows, 8)), block(256);
  if (cols <= 256) launch_pdl(softmax_rows_kernel<8>, grid, block, 0, s, in, rows, cols, scale, out);
  else if (cols <= 1024) launch_pdl(softmax_rows_kernel<32>, grid, block, 0, s, in, rows, cols, scale, out);
  else launch_pdl(softmax_rows_kernel<128>, grid, block, 0, s, in, rows, cols, scale, out);
}

void im2col_input(cudaStream_t s, const float* latents, int n_img, int cin, int h, int w, int kpad, bf16* out,
                  const float* premix) {
  MV_CHECK(kpad >= 9 * cin, "im2col: kpad too small");
  MV_CHECK(kpad % 8 == 0 && (int64_t)n_img * h * w * kpad < (1ll << 31), "im2col: kpad must be a multiple of 8 (32-bit indexing)");
  const int64_t total = (int64_t)n_img * h * w * (kpad / 8);
  launch_pdl(im2col_input_kernel, dim3(grid_for(total, 128)), dim3(128), 0, s, latents, n_img, cin, h, w, kpad, premix, out);
}

// timestep embedding in bf16: the A operand of the time_embedding GEMMs (what autocast feeds linear_1 in the reference)
void timestep_sinusoid_bf16(cudaStream_t s, const int64_t* t, int n, int dim, bf16* out) {
  const int total = n * (dim / 2);
  launch_pdl(sinusoid_kernel<bf16>, dim3(ceil_div(total, 128)), dim3(128), 0, s, t, n, dim, out);
}

size_t groupnorm_scratch_floats(int n_img, int groups) { return (size_t)n_img * GN_MAXP * groups * 2; }

void groupnorm_init() {
  MV_CUDA(cudaFuncSetAttribute(gn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
}

void groupnorm(cudaStream_t s, const bf16* x0, int c0, const bf16* x1, int c1, int n_img, int hw, int groups,
               float eps, const float* gamma, const float* beta, bool silu, bf16* out, float* scratch) {
  const int C = c0 + c1;
  MV_CHECK(C % groups == 0 && groups <= 64, "groupnorm: channels not divisible by groups (<= 64 groups)");
  MV_CHECK(c0 % 8 == 0 && c1 % 8 == 0 && C <= GN_MAXC, "groupnorm: channel counts must be multiples of 8, <= 2560");
  static const int force_two = [] {
    const char* e = getenv("MVLDM_GN_TWO_LAUNCH");
    return e ? atoi(e) : 0;
  }();
  // ---- register-resident path (the model's 320/640/1280/2560-channel tensors at every resolution) ----
  if (!(force_two && n_img >= force_two)) {
    static const bool use_reg = [] {
      const char* e = getenv("MVLDM_GN_REG");
      return !e || atoi(e) != 0;
    }();
    const int cgn = C / groups;
    const int P = (cgn <= 40 && hw >= 512) ? 2 : 1;
    const int tpi = hw % P == 0 ? hw / P : 0;
    const bool pow2 = tpi >= 16 && tpi <= 512 && (tpi & (tpi - 1)) == 0;
    static const bool use_flat = [] {
      const char* e = getenv("MVLDM_GN_FLAT");
      return !e || atoi(e) != 0;
    }();
    const bool flat5 = (cgn == 10 || cgn == 20 || cgn == 40) && C % 40 == 0;
    const bool flat15 = (cgn == 30 || cgn == 60) && C % 120 == 0;   // the 960- / 1920-channel concats of the up path
    const int px_cta = flat5 ? 256 : 128;   // threads x vectors-per-thread / vectors-per-pixel
    if (use_flat && (flat5 || flat15) && hw % px_cta == 0 && hw / px_cta <= 8 && hw >= 256) {
      // one channel block x 256 pixels (320 threads x 4 vectors) or x 128 pixels (480 x 4, the 120-channel blocks) per
      // CTA; larger images are split over a cluster of CTAs.  Small CTAs on purpose: several share an SM, so the load, reduce
      // and store phases of different CTAs overlap (one 640-thread CTA per SM ran at 1.6 TB/s at 64 images)
      const int cs = hw / px_cta, cbk = flat5 ? 40 : 120;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((C / cbk) * n_img * cs, 1, 1);
      cfg.blockDim = dim3(flat5 ? 320 : 480, 1, 1);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = s;
      cudaLaunchAttribute at[2];
      int na = 0;
      if (cs > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cs; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
      }
#ifdef MVLDM_ENABLE_PDL
      if (g_use_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
#endif
      cfg.attrs = at;
      cfg.numAttrs = na;
      const int sl = silu ? 1 : 0;
#define GN_FLAT(NV, RR, CGN) \
  MV_CUDA(cudaLaunchKernelEx(&cfg, gn_flat_kernel<NV, RR, CGN>, x0, c0, x1, c1, hw, eps, gamma, beta, sl, out))
      if (cgn == 10) GN_FLAT(5, 4, 10);
      else if (cgn == 20) GN_FLAT(5, 4, 20);
      else if (cgn == 40) GN_FLAT(5, 4, 40);
      else if (cgn == 30) GN_FLAT(15, 4, 30);
      else GN_FLAT(15, 4, 60);
#undef GN_FLAT
      MV_LAUNCHED();
      return;
    }
    // measured in a CUDA-graph chain (tools/gn_graph_bench.py, 8 images): it wins where a tensor is small (us/launch,
    // this kernel vs the slab kernel: 16 px x 1280 ch 3.9 / 5.9; 64 px x 1280 5.6 / 7.7; 64 px x 640 4.7 / 6.0;
    // 256 px x 640 9.7 / 11.2) and loses at 1024 px (32 / 12: its 80-byte-per-lane strided loads waste sectors)
    const bool wins = hw <= 64 || (hw == 256 && cgn == 20);
    if (use_reg && wins && pow2 && (cgn == 10 || cgn == 20 || cgn == 40 || cgn == 80) && C % (cgn == 80 ? 80 : 40) == 0) {
      const int cbk = cgn == 80 ? 80 : 40;
      const int n_items = (C / cbk) * n_img;
      const int threads = tpi > 64 ? tpi : 64;
      const int ipc = threads / tpi;
      const dim3 grid(ceil_div(n_items, ipc)), block(threads);
      const int sl = silu ? 1 : 0;
#define GN_REG(NV, PP, CGN) \
  launch_pdl(gn_reg_kernel<NV, PP, CGN>, grid, block, 0, s, x0, c0, x1, c1, hw, eps, gamma, beta, sl, tpi, n_items, out)
      if (cgn == 10 && P == 2) GN_REG(5, 2, 10);
      else if (cgn == 10) GN_REG(5, 1, 10);
      else if (cgn == 20 && P == 2) GN_REG(5, 2, 20);
      else if (cgn == 20) GN_REG(5, 1, 20);
      else if (cgn == 40 && P == 2) GN_REG(5, 2, 40);
      else if (cgn == 40) GN_REG(5, 1, 40);
      else GN_REG(10, 1, 80);
#undef GN_REG
      return;
    }
  }
  // ---- single-launch path: one CTA per (image, block of whole groups) ----
  if (!(force_two && n_img >= force_two)) {
    const int cgn = C / groups;
    int cb = cgn;
    while (cb % 8 != 0) cb += cgn;  // lcm(group width, 8)
    if (C % cb == 0 && cb / 8 <= 16 && cb / cgn <= 16 && c0 % 8 == 0) {
      const size_t slab = (size_t)hw * cb * sizeof(bf16);
      const int cache = slab <= 160 * 1024 ? 1 : 0;
      static bool configured[kMaxDevices] = {};
      if (first_use_on_device(configured)) {
        MV_CUDA(cudaFuncSetAttribute(gn_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
      }
      // pixel split over a small cluster when one CTA per (image, channel block) would leave SMs idle
      static const int max_cs = [] {
        const char* e = getenv("MVLDM_GN_CS");
        return e ? atoi(e) : 2;
      }();
      int cs = 1;
      // ... or when its slab (plus 36 KB of reduction scratch) would leave room for only one CTA per SM
      static const size_t slab_limit = [] {
        const char* e = getenv("MVLDM_GN_SLAB_KB");
        return (size_t)(e ? atoi(e) : 64) * 1024;
      }();
      while (cs < max_cs && ((C / cb) * n_img * cs < 148 || slab / cs > slab_limit) && hw % (2 * cs) == 0 &&
             hw / (2 * cs) >= 64)
        cs *= 2;
      const size_t slab_cs = slab / cs;
      const int cache_cs = slab_cs <= 160 * 1024 ? 1 : 0;
      // a slab that does not fit shared memory would be streamed twice with 16-byte accesses at a pitch of C channels (half of
      // every sector wasted): large images (the VAE's 128x128 / 256x256 feature maps) take the coalesced two-launch path below
      if (cache_cs) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((C / cb) * cs, n_img, 1);
      cfg.blockDim = dim3(GNB_THREADS, 1, 1);
      cfg.dynamicSmemBytes = cache_cs ? slab_cs : 0;
      cfg.stream = s;
      cudaLaunchAttribute at[2];
      int na = 0;
      if (cs > 1) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = cs; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
      }
#ifdef MVLDM_ENABLE_PDL
      if (g_use_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
      }
#endif
      cfg.attrs = at;
      cfg.numAttrs = na;
      MV_CUDA(cudaLaunchKernelEx(&cfg, gn_block_kernel, x0, c0, x1, c1, hw, groups, eps, gamma, beta, silu ? 1 : 0, cb,
                                 cache_cs, out));
      MV_LAUNCHED();
      (void)cache;
      return;
      }
    }
  }
  // ---- two-launch path (large images / odd sizes) ----
  int P = 1;
  while (P < GN_MAXP && n_img * P < 592 && hw % (2 * P) == 0 && hw / (2 * P) >= 4) P *= 2;
  launch_pdl(gn_partial_kernel, dim3(P, n_img), dim3(256), 0, s, x0, c0, x1, c1, hw, groups, hw / P, scratch);
  launch_pdl(gn_apply_kernel, dim3(P, n_img), dim3(256), 0, s, x0, c0, x1, c1, hw, groups, eps, P, (const float*)scratch,
             gamma, beta, silu ? 1 : 0, out);
}

void layernorm(cudaStream_t s, const bf16* x, int rows, int c, float eps, const float* gamma, const float* beta,
               bf16* out) {
  MV_CHECK(c % 8 == 0 && c <= 32 * 8 * 8, "layernorm: unsupported channel count");
  const int warps = 4;
  const dim3 grid(ceil_div(rows, warps)), block(warps * 32);
  if (c <= 32 * 8 * 2) launch_pdl(layernorm_kernel<2>, grid, block, 0, s, x, rows, c, eps, gamma, beta, out);
  else if (c <= 32 * 8 * 5) launch_pdl(layernorm_kernel<5>, grid, block, 0, s, x, rows, c, eps, gamma, beta, out);
  else launch_pdl(layernorm_kernel<8>, grid, block, 0, s, x, rows, c, eps, gamma, beta, out);
}

void upsample_nearest2x(cudaStream_t s, const bf16* x, int n_img, int h, int w, int c, bf16* out) {
  MV_CHECK(c % 8 == 0, "upsample: channels must be a multiple of 8");
  const int64_t total = (int64_t)n_img * 4 * h * w * (c / 8);
  launch_pdl(upsample2x_kernel, dim3(grid_for(total, 256)), dim3(256), 0, s, x, n_img, h, w, c, out);
}

}  // namespace mvldm
