// Developer harness: floor cost of a dependent kernel chain inside a CUDA graph, with and without PDL, vs a grid barrier.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
__global__ void __launch_bounds__(320, 1) tiny(float* p, int pdl) {
  extern __shared__ float sm[];
  if (pdl) { asm volatile("griddepcontrol.wait;" ::: "memory"); asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
  if (threadIdx.x == 0) p[blockIdx.x] += 1.0f;
}
__global__ void __launch_bounds__(320, 1) barrier_chain(float* p, unsigned* counter, int phases) {
  extern __shared__ float sm[];
  for (int ph = 0; ph < phases; ++ph) {
    if (threadIdx.x == 0) p[blockIdx.x] += 1.0f;
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicAdd(counter, 1u);
      const unsigned target = (unsigned)(ph + 1) * gridDim.x;
      unsigned v;
      do { asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter)); } while (v < target);
      __threadfence();
    }
    __syncthreads();
  }
}
static float run_graph(int n, int grid, size_t smem, int pdl, float* d) {
  cudaStream_t st; CK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CK(cudaFuncSetAttribute(tiny, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int i = 0; i < n; ++i) {
    cudaLaunchConfig_t cfg{}; cfg.gridDim = grid; cfg.blockDim = 320; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, tiny, d, pdl));
  }
  CK(cudaStreamEndCapture(st, &g)); CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0, best = 1e9;
  for (int rep = 0; rep < 5; ++rep) { CK(cudaEventRecord(e0, st)); CK(cudaGraphLaunch(ge, st)); CK(cudaEventRecord(e1, st)); CK(cudaStreamSynchronize(st)); CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms; }
  return best * 1000.f / n;
}
int main() {
  float* d; CK(cudaMalloc(&d, 4096)); CK(cudaMemset(d, 0, 4096));
  for (int grid : {1, 100, 148}) for (size_t smem : {(size_t)0, (size_t)197 * 1024}) for (int pdl : {0, 1})
    printf("graph chain of 300 kernels: grid=%3d smem=%6zu pdl=%d : %.2f us per kernel\n", grid, smem, pdl, run_graph(300, grid, smem, pdl, d));
  unsigned* c; CK(cudaMalloc(&c, 4));
  CK(cudaFuncSetAttribute(barrier_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(c, 0, 4));
    CK(cudaEventRecord(e0)); barrier_chain<<<148, 320, 197 * 1024>>>(d, c, 300); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    printf("persistent kernel, 300 grid barriers over 148 CTAs: %.2f us per phase\n", ms * 1000.f / 300);
  }
  return 0;
}
