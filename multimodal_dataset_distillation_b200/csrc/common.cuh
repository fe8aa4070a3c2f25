// Shared device/host helpers for the vldd_b200 library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#define VLDD_OK 0
#ifndef VLDD_ERR_ARG
#define VLDD_ERR_ARG (-1)
#define VLDD_ERR_CUDA (-2)
#define VLDD_ERR_WORKSPACE (-3)
#endif

namespace vldd {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define VLDD_REQUIRE(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      vldd::set_error(__VA_ARGS__);             \
      return VLDD_ERR_ARG;                      \
    }                                           \
  } while (0)

#define VLDD_CUDA(call)                                                        \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      vldd::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
      return VLDD_ERR_CUDA;                                                    \
    }                                                                          \
  } while (0)

constexpr int kMaxSMs = 256;   // sizing bound for per-block scratch (B200: 148)
// SM count of the current device, queried once per device ordinal (grids are sized in multiples of it)
inline int num_sms() {
  static int cached[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0 || n > kMaxSMs) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every kernel of the library starts with pdl_wait(): it blocks until the preceding kernel in the stream has
// completed and its writes are visible (a no-op when the launch carries no programmatic dependency), then
// pdl_launch_dependents() lets the NEXT kernel be scheduled early so that its launch latency and prologue
// (barrier init, TMEM allocation, descriptor prefetch) overlap this kernel's execution.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_launch_dependents(); }

bool pdl_enabled();   // VLDD_PDL=0 disables the launch attribute (capi.cu)

// Number of kernels of this library that have been (or, while a CUDA graph is being captured, will be per replay)
// launched by the process: launch_k() counts, the engine adds a graph's node count on every replay.  bench.py reports
// the difference across its timed region as `gpu_launches` (vldd_kernel_launch_count).
inline std::atomic<unsigned long long>& launch_counter() {
  static std::atomic<unsigned long long> c{0};
  return c;
}

template <typename... KArgs, typename... Args>
inline void launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);   // errors surface through cudaGetLastError (check_launch)
  launch_counter().fetch_add(1, std::memory_order_relaxed);
}

__host__ __device__ inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- warp / block reductions (all threads of the block must call) -------------------------------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum, result broadcast to every thread. `scratch` holds >= 33 elements of T.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse across consecutive calls
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    T t = lane < nw ? scratch[lane] : T(0);
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    float t = lane < nw ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// ---- streaming loads/stores: 128-bit, bypass L1 allocation for touch-once data -------------------
__device__ __forceinline__ float4 ldg_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void stg_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

// ---- exact-erf GELU pieces (networks.py:634 nn.GELU() default) -----------------------------------
__device__ __forceinline__ void gelu_parts(float p, float& phi, float& dphi, float& ddphi) {
  const float cdf = 0.5f * (1.0f + erff(p * 0.70710678118654752440f));
  const float pdf = expf(-0.5f * p * p) * 0.39894228040143267794f;
  phi = p * cdf;
  dphi = cdf + p * pdf;
  ddphi = pdf * (2.0f - p * p);
}

}  // namespace vldd
