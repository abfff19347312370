// Micro-benchmark: per-kernel cost of a dependent chain of short kernels, with and without programmatic dependent launch,
// as plain stream launches and as a CUDA graph.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_chain pdl_chain.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void spin_ns(long long ns) {
  long long t0; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
  long long t = t0;
  while (t - t0 < ns) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
}

// mode bit0: griddepcontrol.wait after the prologue; bit1: launch_dependents early (after wait), bit2: launch_dependents late
__global__ void work(const float* in, float* out, int n, int prologue_ns, int body_ns, int tail_ns, int mode) {
  extern __shared__ float sm[];
  spin_ns(prologue_ns);                       // independent prologue (barrier init, TMEM alloc, descriptor prefetch ...)
  if (mode & 1) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (mode & 2) asm volatile("griddepcontrol.launch_dependents;");
  float acc = 0.f;
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += gridDim.x * blockDim.x) acc += in[i];
  spin_ns(body_ns);
  if (mode & 4) asm volatile("griddepcontrol.launch_dependents;");
  spin_ns(tail_ns);                           // epilogue
  for (int i = threadIdx.x + blockIdx.x * blockDim.x; i < n; i += gridDim.x * blockDim.x) out[i] = acc * 1e-9f + in[i];
  if (threadIdx.x == 0 && n < 0) sm[0] = acc;
}

static float run(bool graph, bool pdl, int mode, int smem, int chain, int pro, int body, int tail, float* a, float* b, int n) {
  cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  CK(cudaFuncSetAttribute(work, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  auto enqueue = [&]() {
    for (int i = 0; i < chain; ++i) {
      cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(148); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem; cfg.stream = s;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
      const float* in = (i & 1) ? b : a; float* out = (i & 1) ? a : b;
      CK(cudaLaunchKernelEx(&cfg, work, in, out, n, pro, body, tail, pdl ? mode : 0));
    }
  };
  cudaGraphExec_t ge = nullptr;
  if (graph) {
    cudaGraph_t g; CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal)); enqueue(); CK(cudaStreamEndCapture(s, &g));
    CK(cudaGraphInstantiate(&ge, g, 0)); cudaGraphDestroy(g);
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    CK(cudaStreamSynchronize(s));
    CK(cudaEventRecord(e0, s));
    if (graph) CK(cudaGraphLaunch(ge, s)); else enqueue();
    CK(cudaEventRecord(e1, s));
    CK(cudaStreamSynchronize(s));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep && ms < best) best = ms;
  }
  if (ge) cudaGraphExecDestroy(ge);
  cudaStreamDestroy(s);
  return best * 1000.f / chain;
}

int main() {
  const int n = 148 * 256 * 4, chain = 200;
  float *a, *b; CK(cudaMalloc(&a, n * 4)); CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(a, 0, n * 4)); CK(cudaMemset(b, 0, n * 4));
  printf("per-kernel us in a %d-long dependent chain (148 CTAs x 256 thr)\n", chain);
  printf("%-44s %8s %8s %8s %8s %8s\n", "prologue/body/tail ns, smem", "stream", "graph", "g+pdl:2", "g+pdl:4", "s+pdl:2");
  const int cases[][4] = {{0, 0, 0, 0}, {1000, 3000, 1000, 0}, {1000, 3000, 1000, 100 * 1024}, {1000, 3000, 1000, 200 * 1024},
                          {1000, 8000, 1000, 100 * 1024}, {1000, 8000, 1000, 200 * 1024}};
  for (auto& c : cases) {
    float t0 = run(false, false, 0, c[3], chain, c[0], c[1], c[2], a, b, n);
    float t1 = run(true, false, 0, c[3], chain, c[0], c[1], c[2], a, b, n);
    float t2 = run(true, true, 1 | 2, c[3], chain, c[0], c[1], c[2], a, b, n);
    float t3 = run(true, true, 1 | 4, c[3], chain, c[0], c[1], c[2], a, b, n);
    float t4 = run(false, true, 1 | 2, c[3], chain, c[0], c[1], c[2], a, b, n);
    char lab[64]; snprintf(lab, 64, "%d/%d/%d ns, %d KB", c[0], c[1], c[2], c[3] / 1024);
    printf("%-44s %8.2f %8.2f %8.2f %8.2f %8.2f\n", lab, t0, t1, t2, t3, t4);
  }
  return 0;
}
