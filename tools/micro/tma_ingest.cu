// Microbenchmark: how fast can one SM pull TMA tiles (SWIZZLE_128B, 128-byte rows) from L2 / HBM?
// grid = #SMs, each CTA streams `iters` stages of (rowsA + rowsB) x 128 B through a ring of smem stages; a consumer
// thread just releases the stage.  Reports GB/s per SM and aggregate.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include "../../mvldm_b200/csrc/tc_common.cuh"
using namespace mvldm;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

struct P { CUtensorMap tm; int iters; int rows; int nbox; int issuers; uint64_t span_rows; };

template <int STAGES>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ P p) {
  extern __shared__ uint8_t raw[];
  __shared__ __align__(8) uint64_t full[STAGES], empty[STAGES];
  const uint32_t base = (tc::smem_u32(raw) + 1023u) & ~1023u;
  const int stage_bytes = p.rows * 128 * p.nbox;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { tc::mbar_init(tc::smem_u32(&full[s]), 1); tc::mbar_init(tc::smem_u32(&empty[s]), 1); }
    tc::mbar_fence_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < p.iters; ++i) {
      const int s = i % STAGES;
      tc::mbar_wait(tc::smem_u32(&empty[s]), ((i / STAGES) & 1) ^ 1);
      tc::mbar_expect_tx(tc::smem_u32(&full[s]), stage_bytes);
      for (int b = 0; b < p.nbox; ++b) {
        const uint64_t row = ((uint64_t)(blockIdx.x * 7919 + i * 131 + b * 17) * p.rows) % (p.span_rows - p.rows);
        tc::tma_load_2d(base + s * stage_bytes + b * p.rows * 128, &p.tm, tc::smem_u32(&full[s]), (int)((i * 64) % 4096), (int)row);
      }
    }
  } else if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.iters; ++i) {
      const int s = i % STAGES;
      tc::mbar_wait(tc::smem_u32(&full[s]), (i / STAGES) & 1);
      tc::mbar_arrive(tc::smem_u32(&empty[s]));
    }
  }
}

int main() {
  PFN_cuTensorMapEncodeTiled_v12000 enc; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const uint64_t cols = 4096 + 64;           // bf16 per row (row pitch 8320 B)
  for (int big = 0; big < 2; ++big) {
    const uint64_t rows = big ? 400000 : 12000;   // 3.3 GB (HBM) vs 100 MB (L2-resident)
    void* buf; CK(cudaMalloc(&buf, rows * cols * 2)); CK(cudaMemset(buf, 1, rows * cols * 2));
    for (int boxrows : {128, 256}) for (int nbox : {1, 2}) {
      CUtensorMap tm; cuuint64_t dims[2] = {cols, rows}; cuuint64_t str[1] = {cols * 2}; cuuint32_t box[2] = {64, (cuuint32_t)boxrows}; cuuint32_t es[2] = {1, 1};
      if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
      for (int grid : {148, 74, 16}) {
        P p{tm, 400, boxrows, nbox, 1, rows};
        const int smem = 4 * boxrows * 128 * nbox + 1024;
        CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<4><<<grid, 128, smem>>>(p); CK(cudaDeviceSynchronize());
        cudaEventRecord(e0); k<4><<<grid, 128, smem>>>(p); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double bytes = (double)grid * p.iters * boxrows * 128.0 * nbox;
        printf("%s box %dx128B x%d, 4 stages (%d KB in flight), grid %3d: %7.1f GB/s per SM, %6.2f TB/s total\n", big ? "HBM" : "L2 ", boxrows, nbox,
               4 * boxrows * nbox / 8, grid, bytes / ms * 1e-6 / grid, bytes / ms * 1e-9);
      }
    }
    cudaFree(buf);
  }
  return 0;
}
