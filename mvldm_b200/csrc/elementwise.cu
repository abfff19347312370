// HBM/L2-bound helper kernels of the MV-LDM hot path: input concat/im2col, timestep embedding,
// GroupNorm(+SiLU), LayerNorm, nearest upsample, CFG+DDIM update, ray maps, weight packing.
// All activations are bf16 NHWC ([image, h, w, channel]); statistics and scalars are fp32.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace mvldm {

thread_local int g_launch_count = 0;
bool g_use_pdl = []() {  // opt-in: measured 2% slower than plain stream order inside the captured graph (profiles/)
  const char* e = getenv("MVLDM_PDL");
  return e && e[0] == '1';
}();

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

__global__ void nhwc_to_nchw_kernel(const bf16* __restrict__ x, int n_img, int hw, int c, float* __restrict__ out) {
  const int64_t total = (int64_t)n_img * hw * c;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw);
    const int ch = (int)((i / hw) % c);
    const int img = (int)(i / ((int64_t)hw * c));
    out[i] = __bfloat162float(x[((int64_t)img * hw + p) * c + ch]);
  }
}

// inputs[b, v, :, p] = [latent(4) | mask | rays(R)], context views first (diffusion_wrapper.py:429-432)
__global__ void build_inputs_kernel(const float* __restrict__ x_t, const float* __restrict__ ctx,
                                    const float* __restrict__ rays, int B, int v_c, int v_t, int ray_views, int ray_off,
                                    int R, int hw, float* __restrict__ out) {
  const int V = v_c + v_t, C = 5 + R;
  const int64_t total = (int64_t)B * V * C * hw;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int p = (int)(i % hw);
    const int ch = (int)((i / hw) % C);
    const int v = (int)((i / ((int64_t)hw * C)) % V);
    const int b = (int)(i / ((int64_t)hw * C * V));
    float val;
    if (ch < 4)
      val = v < v_c ? ctx[(((int64_t)b * v_c + v) * 4 + ch) * hw + p] : x_t[(((int64_t)b * v_t + (v - v_c)) * 4 + ch) * hw + p];
    else if (ch == 4)
      val = v < v_c ? 0.f : 1.f;
    else
      val = rays[(((int64_t)b * ray_views + ray_off + v) * R + (ch - 5)) * hw + p];
    out[i] = val;
  }
}

// CFG compose + DDIM update (eta = 0, epsilon prediction), K14 + K15
__global__ void ddim_step_kernel(const float* __restrict__ eps_c, const float* __restrict__ eps_u, float scale, int B,
                                 int v_c, int v_t, int chw, const float* __restrict__ x_t, float sa, float s1a, float sp,
                                 float s1p, float* __restrict__ x_prev, float* __restrict__ eps_out) {
  const int64_t per_scene = (int64_t)v_t * chw;
  const int64_t total = (int64_t)B * per_scene;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_scene, r = i - b * per_scene;
    float e = eps_c[(b * (v_c + v_t) + v_c) * chw + r];
    if (eps_u) {
      const float u = eps_u[i];
      e = u + scale * (e - u);
    }
    const float x = x_t[i];
    const float x0 = (x - s1a * e) / sa;
    x_prev[i] = sp * x0 + s1p * e;
    if (eps_out) eps_out[i] = e;
  }
}

// CFG compose + DDPM ancestral update (diffusers DDPMScheduler.step, epsilon prediction): x0 = (x - sqrt(1-a_t) e) / sqrt(a_t),
// optionally clamped to +-clip, x_prev = c_x0 x0 + c_xt x + sigma z
__global__ void ddpm_step_kernel(const float* __restrict__ eps_c, const float* __restrict__ eps_u, float scale, int B,
                                 int v_c, int v_t, int chw, const float* __restrict__ x_t, const float* __restrict__ noise,
                                 float sa, float s1a, float c_x0, float c_xt, float sigma, float clip,
                                 float* __restrict__ x_prev) {
  const int64_t per_scene = (int64_t)v_t * chw;
  const int64_t total = (int64_t)B * per_scene;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = i / per_scene, r = i - b * per_scene;
    float e = eps_c[(b * (v_c + v_t) + v_c) * chw + r];
    if (eps_u) {
      const float u = eps_u[i];
      e = u + scale * (e - u);
    }
    const float x = x_t[i];
    float x0 = (x - s1a * e) / sa;
    if (clip > 0.f) x0 = fminf(fmaxf(x0, -clip), clip);
    float y = c_x0 * x0 + c_xt * x;
    if (noise) y = fmaf(sigma, noise[i], y);
    x_prev[i] = y;
  }
}

// K17: pixel-centre grid -> K^-1 -> normalise -> rotate; origin broadcast; optional Pluecker moment
// oct_o / oct_d > 0: the origin / direction triple is replaced by its PositionalEncoding (use_ray_encoding: true;
// src/model/encodings/positional_encoding.py:28-49): channel (d f p) = sin(2 pi 2^f x_d + p pi/2), d-major, then f, then p
__device__ __forceinline__ int write_encoded(float* o, int s, int ch, const float v[3], int oct, int srt) {
  if (oct <= 0) {
    for (int d = 0; d < 3; ++d) o[(int64_t)(ch + d) * s] = v[d];
    return ch + 3;
  }
  if (srt) {  // srt_ray_encoding: true - src/model/srt/layers.py:9-32: [sin(pi 2^f x_d) over (d f) | cos(...) over (d f)]
    for (int d = 0; d < 3; ++d)
      for (int f = 0; f < oct; ++f) {
        const float arg = v[d] * (exp2f((float)f) * 3.141592653589793f);
        o[(int64_t)(ch + d * oct + f) * s] = sinf(arg);
        o[(int64_t)(ch + 3 * oct + d * oct + f) * s] = cosf(arg);
      }
    return ch + 6 * oct;
  }
  for (int d = 0; d < 3; ++d)
    for (int f = 0; f < oct; ++f) {
      const float freq = 6.283185307179586f * exp2f((float)f);   // fp32(2 pi) * 2^f, as torch builds it
      const float arg = v[d] * freq;
      o[(int64_t)(ch++) * s] = sinf(arg + 0.f);
      o[(int64_t)(ch++) * s] = sinf(arg + 1.5707963267948966f);
    }
  return ch;
}

__global__ void raymap_kernel(const float* __restrict__ extr, const float* __restrict__ intr, int n, int h, int w,
                              int plucker, int oct_o, int oct_d, int srt, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * h * w) return;
  const int p = i % (h * w), v = i / (h * w);
  const float* K = intr + v * 9;
  const float* E = extr + v * 16;
  const float a = K[0], b = K[1], c = K[2], d = K[3], e = K[4], f = K[5], g = K[6], hh = K[7], k = K[8];
  const float A = e * k - f * hh, Bc = -(d * k - f * g), Cc = d * hh - e * g;
  const float det = a * A + b * Bc + c * Cc;
  const float inv[9] = {A / det,  -(b * k - c * hh) / det, (b * f - c * e) / det,
                        Bc / det, (a * k - c * g) / det,   -(a * f - c * d) / det,
                        Cc / det, -(a * hh - b * g) / det, (a * e - b * d) / det};
  const float px = ((float)(p % w) + 0.5f) / (float)w, py = ((float)(p / w) + 0.5f) / (float)h;
  float dx = inv[0] * px + inv[1] * py + inv[2];
  float dy = inv[3] * px + inv[4] * py + inv[5];
  float dz = inv[6] * px + inv[7] * py + inv[8];
  const float nrm = sqrtf(dx * dx + dy * dy + dz * dz);
  dx /= nrm; dy /= nrm; dz /= nrm;
  const float wx = E[0] * dx + E[1] * dy + E[2] * dz;
  const float wy = E[4] * dx + E[5] * dy + E[6] * dz;
  const float wz = E[8] * dx + E[9] * dy + E[10] * dz;
  float ox = E[3], oy = E[7], oz = E[11];
  if (plucker) {
    const float mx = oy * wz - oz * wy, my = oz * wx - ox * wz, mz = ox * wy - oy * wx;
    ox = mx; oy = my; oz = mz;
  }
  const int s = h * w;
  const int C = (oct_o > 0 ? 6 * oct_o : 3) + (oct_d > 0 ? 6 * oct_d : 3);
  float* o = out + (int64_t)v * C * s + p;
  const float ov[3] = {ox, oy, oz}, dv[3] = {wx, wy, wz};
  const int ch = write_encoded(o, s, 0, ov, oct_o, srt);
  write_encoded(o, s, ch, dv, oct_d, srt);
}

__global__ void convert_kernel(const void* __restrict__ src, int dtype, int64_t n, float* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v;
    if (dtype == MVLDM_F32) v = reinterpret_cast<const float*>(src)[i];
    else if (dtype == MVLDM_BF16) v = __bfloat162float(reinterpret_cast<const bf16*>(src)[i]);
    else v = __half2float(reinterpret_cast<const __half*>(src)[i]);
    dst[i] = v;
  }
}

// dst[rowmap ? rowmap[r] : r, dst_col0 + j] = src[r, j]
__global__ void pack_rows_kernel(const float* __restrict__ src, int rows, int cols, int src_ld, bf16* __restrict__ dst,
                                 int dst_ld, int dst_col0, const int* __restrict__ rowmap) {
  const int64_t total = (int64_t)rows * cols;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), j = (int)(i % cols);
    const int dr = rowmap ? rowmap[r] : r;
    dst[(int64_t)dr * dst_ld + dst_col0 + j] = __float2bfloat16(src[(int64_t)r * src_ld + j]);
  }
}

// conv filter [cout, cin, ks, ks] -> dst[cout, dst_col0 + tap*cin + c]  (tap-major K, matches the implicit-GEMM A order)
__global__ void pack_conv_kernel(const float* __restrict__ w, int cout, int cin, int ks, bf16* __restrict__ dst,
                                 int dst_ld, int dst_col0) {
  const int taps = ks * ks;
  const int64_t total = (int64_t)cout * cin * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % cin);
    const int tap = (int)((i / cin) % taps);
    const int o = (int)(i / ((int64_t)cin * taps));
    dst[(int64_t)o * dst_ld + dst_col0 + tap * cin + c] = __float2bfloat16(w[((int64_t)o * cin + c) * taps + tap]);
  }
}

inline int grid_for(int64_t total, int threads) {
  int64_t b = (total + threads - 1) / threads;
  const int64_t cap = 148 * 16;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

void nhwc_to_nchw_f32(cudaStream_t s, const bf16* x, int n_img, int hw, int c, float* out) {
  const int64_t total = (int64_t)n_img * hw * c;
  nhwc_to_nchw_kernel<<<grid_for(total, 256), 256, 0, s>>>(x, n_img, hw, c, out);
  MV_LAUNCHED();
}

void build_inputs(cudaStream_t s, const float* x_t, const float* ctx, const float* rays, int B, int v_c, int v_t,
                  int ray_views, int ray_off, int R, int hw, float* out) {
  const int64_t total = (int64_t)B * (v_c + v_t) * (5 + R) * hw;
  build_inputs_kernel<<<grid_for(total, 256), 256, 0, s>>>(x_t, ctx, rays, B, v_c, v_t, ray_views, ray_off, R, hw, out);
  MV_LAUNCHED();
}

void ddim_step(cudaStream_t s, const float* eps_c, const float* eps_u, float scale, int B, int v_c, int v_t, int chw,
               const float* x_t, float sa, float s1a, float sp, float s1p, float* x_prev, float* eps_out) {
  const int64_t total = (int64_t)B * v_t * chw;
  ddim_step_kernel<<<grid_for(total, 256), 256, 0, s>>>(eps_c, eps_u, scale, B, v_c, v_t, chw, x_t, sa, s1a, sp, s1p,
                                                         x_prev, eps_out);
  MV_LAUNCHED();
}

void ddpm_step(cudaStream_t s, const float* eps_c, const float* eps_u, float scale, int B, int v_c, int v_t, int chw,
               const float* x_t, const float* noise, float sa, float s1a, float c_x0, float c_xt, float sigma, float clip,
               float* x_prev) {
  const int64_t total = (int64_t)B * v_t * chw;
  ddpm_step_kernel<<<grid_for(total, 256), 256, 0, s>>>(eps_c, eps_u, scale, B, v_c, v_t, chw, x_t, noise, sa, s1a, c_x0, c_xt,
                                                         sigma, clip, x_prev);
  MV_LAUNCHED();
}

void raymap(cudaStream_t s, const float* extr, const float* intr, int n, int h, int w, bool plucker, float* out, int oct_o,
            int oct_d, bool srt) {
  raymap_kernel<<<ceil_div(n * h * w, 128), 128, 0, s>>>(extr, intr, n, h, w, plucker ? 1 : 0, oct_o, oct_d, srt ? 1 : 0, out);
  MV_LAUNCHED();
}

void convert_f32(cudaStream_t s, const void* src, int dtype, int64_t n, float* dst) {
  convert_kernel<<<grid_for(n, 256), 256, 0, s>>>(src, dtype, n, dst);
  MV_LAUNCHED();
}

void pack_rows(cudaStream_t s, const float* src, int rows, int cols, int src_ld, bf16* dst, int dst_ld, int dst_col0,
               const int* rowmap) {
  pack_rows_kernel<<<grid_for((int64_t)rows * cols, 256), 256, 0, s>>>(src, rows, cols, src_ld, dst, dst_ld, dst_col0,
                                                                       rowmap);
  MV_LAUNCHED();
}

void pack_conv3x3(cudaStream_t s, const float* w, int cout, int cin, int ks, bf16* dst, int dst_ld, int dst_col0) {
  pack_conv_kernel<<<grid_for((int64_t)cout * cin * ks * ks, 256), 256, 0, s>>>(w, cout, cin, ks, dst, dst_ld, dst_col0);
  MV_LAUNCHED();
}

// pack-time only: C[m, n] = A[m, k] . B[k, n] (+ bias_in[k] folded as C_bias[m] = A . bias_in + bias_add), plain fp32, row-major
__global__ void matmul_f32_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, int m, int n,
                                  int k) {
  __shared__ float sa[16][17], sb[16][17];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int row = blockIdx.y * 16 + ty, col = blockIdx.x * 16 + tx;
  float acc = 0.f;
  for (int k0 = 0; k0 < k; k0 += 16) {
    sa[ty][tx] = (row < m && k0 + tx < k) ? A[(int64_t)row * k + k0 + tx] : 0.f;
    sb[ty][tx] = (k0 + ty < k && col < n) ? B[(int64_t)(k0 + ty) * n + col] : 0.f;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 16; ++j) acc = fmaf(sa[ty][j], sb[j][tx], acc);
    __syncthreads();
  }
  if (row < m && col < n) C[(int64_t)row * n + col] = acc;
}
__global__ void matvec_bias_kernel(const float* __restrict__ A, const float* __restrict__ x, const float* __restrict__ add,
                                   float* __restrict__ y, int m, int k) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= m) return;
  float acc = add ? add[row] : 0.f;
  for (int j = 0; j < k; ++j) acc = fmaf(A[(int64_t)row * k + j], x[j], acc);
  y[row] = acc;
}
void matmul_f32(cudaStream_t s, const float* A, const float* B, float* C, int m, int n, int k) {
  matmul_f32_kernel<<<dim3(ceil_div(n, 16), ceil_div(m, 16)), dim3(16, 16), 0, s>>>(A, B, C, m, n, k);
  MV_LAUNCHED();
}
void matvec_bias(cudaStream_t s, const float* A, const float* x, const float* add, float* y, int m, int k) {
  matvec_bias_kernel<<<ceil_div(m, 128), 128, 0, s>>>(A, x, add, y, m, k);
  MV_LAUNCHED();
}

void fill_zero(cudaStream_t s, void* p, size_t bytes) { MV_CUDA(cudaMemsetAsync(p, 0, bytes, s)); }

}  // namespace mvldm
