// MUFU.EX2 throughput per SM: f32 vs packed bf16x2 / f16x2 (two results per instruction).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu mufu.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void k(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
float run(uint32_t* out, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 1024>>>(out, 10, 1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 2, 1024>>>(out, iters, 1);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  uint32_t* out; CK(cudaMalloc(&out, 148 * 2 * 1024 * 4));
  const int iters = 4000;
  const double instr = 148.0 * 2 * 1024 * 8.0 * iters;  // thread-instructions
  const char* names[3] = {"ex2.approx.ftz.f32", "ex2.approx.ftz.bf16x2", "ex2.approx.f16x2"};
  float ms[3] = {run<0>(out, iters), run<1>(out, iters), run<2>(out, iters)};
  CK(cudaGetLastError());
  for (int m = 0; m < 3; ++m)
    printf("%-24s %.3f ms  %.1f thread-instr/clk/SM (at 1.9 GHz)  %.2f T results/s\n", names[m], ms[m],
           instr / (ms[m] * 1e-3) / 148 / 1.9e9, instr * (m ? 2 : 1) / (ms[m] * 1e-3) / 1e12);
  return 0;
}
