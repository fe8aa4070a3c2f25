// Developer harness (not part of the library): structure of a "2 reads + 1 write + block reduction" streaming kernel
// (the matching-loss numerator + adjoint pass) against the plain 2R+1W kernel, on operands larger than L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I. -I../../include dev/stream_bench_test.cu -o dev/stream_bench_test
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../common.cuh"

namespace vldd {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
bool pdl_enabled() { return true; }
int check_launch(const char*) { return 0; }
}  // namespace vldd
using namespace vldd;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

// MODE bits: 1 = reduce, 2 = four independent accumulators, 4 = streaming store, 8 = unroll 4 (else 2), 16 = fp32 block reduce
template <int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS) pass_kernel(const float* __restrict__ x, const float* __restrict__ t, float c, int64_t n4,
                                                       double* __restrict__ parts, float* __restrict__ a) {
  pdl_enter();
  __shared__ double scratch[34];
  __shared__ float fscratch[34];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (int64_t)gridDim.x * blockDim.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int U = (MODE & 8) ? 4 : 2;
  auto body = [&](const float4 xv, const float4 tv, int64_t i) {
    const float4 d = make_float4(xv.x - tv.x, xv.y - tv.y, xv.z - tv.z, xv.w - tv.w);
    if (MODE & 1) {
      if (MODE & 2) { acc[0] = fmaf(d.x, d.x, acc[0]); acc[1] = fmaf(d.y, d.y, acc[1]); acc[2] = fmaf(d.z, d.z, acc[2]); acc[3] = fmaf(d.w, d.w, acc[3]); }
      else { acc[0] = fmaf(d.x, d.x, acc[0]); acc[0] = fmaf(d.y, d.y, acc[0]); acc[0] = fmaf(d.z, d.z, acc[0]); acc[0] = fmaf(d.w, d.w, acc[0]); }
    }
    const float4 o = make_float4(c * d.x, c * d.y, c * d.z, c * d.w);
    if (MODE & 4) stg_stream4(a + 4 * i, o); else *reinterpret_cast<float4*>(a + 4 * i) = o;
  };
  int64_t i = tid;
  for (; i + (U - 1) * stride < n4; i += U * stride) {
    float4 xv[U], tv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) { xv[u] = ldg_stream4(x + 4 * (i + u * stride)); tv[u] = ldg_stream4(t + 4 * (i + u * stride)); }
#pragma unroll
    for (int u = 0; u < U; ++u) body(xv[u], tv[u], i + u * stride);
  }
  for (; i < n4; i += stride) body(ldg_stream4(x + 4 * i), ldg_stream4(t + 4 * i), i);
  if (MODE & 1) {
    const float s = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    if (MODE & 16) {
      const float b = block_sum<float>(s, fscratch);
      if (threadIdx.x == 0) parts[blockIdx.x] = (double)b;
    } else {
      const double b = block_sum<double>((double)s, scratch);
      if (threadIdx.x == 0) parts[blockIdx.x] = b;
    }
  }
}

template <int MODE, int THREADS>
static void run(const char* name, int blocks_per_sm, std::vector<float*>& bufs, int64_t n, double* parts) {
  const int nb = (int)bufs.size();
  const int grid = 148 * blocks_per_sm;
  auto one = [&](int i) {
    // disjoint (x, t, a) triples: a buffer comes back only after every other one has been touched (footprint > L2)
    const int k = (3 * i) % nb;
    launch_k(pass_kernel<MODE, THREADS>, grid, THREADS, 0, (cudaStream_t)0, (const float*)bufs[k], (const float*)bufs[k + 1], 0.5f, n / 4,
             parts, bufs[k + 2]);
  };
  for (int i = 0; i < nb; ++i) one(i);
  CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 3 * nb; ++i) one(i + rep);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  const double us = best * 1000 / (3 * nb);
  printf("%-44s threads=%4d blocks/SM=%d  %6.2f us  %7.1f GB/s\n", name, THREADS, blocks_per_sm, us, 12.0 * n / (us * 1e-6) / 1e9);
  fflush(stdout);
}

int main() {
  const int64_t n = 7087104;
  std::vector<float*> bufs(18);          // 6 triples x 85 MB = 510 MB
  for (auto& b : bufs) { CK(cudaMalloc(&b, n * 4)); CK(cudaMemset(b, 0, n * 4)); }
  double* parts; CK(cudaMalloc(&parts, 8 * 4096));
  run<0, 256>("2R+1W, no reduction", 8, bufs, n, parts);
  run<4, 256>("2R+1W, no reduction, streaming store", 8, bufs, n, parts);
  run<1, 256>("reduce: 1 acc, fp64 block sum (current)", 8, bufs, n, parts);
  run<1 | 4, 256>("reduce: 1 acc, fp64, streaming store", 8, bufs, n, parts);
  run<1 | 2, 256>("reduce: 4 acc, fp64 block sum", 8, bufs, n, parts);
  run<1 | 2 | 16, 256>("reduce: 4 acc, fp32 block sum", 8, bufs, n, parts);
  run<1 | 2 | 4 | 16, 256>("reduce: 4 acc, fp32, streaming store", 8, bufs, n, parts);
  run<1 | 2 | 4 | 8 | 16, 256>("reduce: 4 acc, fp32, stream st, unroll 4", 8, bufs, n, parts);
  run<1 | 2 | 4 | 16, 256>("reduce: 4 acc, fp32, stream st", 4, bufs, n, parts);
  run<1 | 2 | 4 | 16, 512>("reduce: 4 acc, fp32, stream st", 4, bufs, n, parts);
  run<1 | 2 | 4 | 16, 512>("reduce: 4 acc, fp32, stream st", 2, bufs, n, parts);
  run<1 | 2 | 4 | 8 | 16, 512>("reduce: 4 acc, fp32, stream st, unroll 4", 2, bufs, n, parts);
  run<1 | 2 | 4 | 16, 1024>("reduce: 4 acc, fp32, stream st", 2, bufs, n, parts);
  run<1 | 2 | 4 | 16, 1024>("reduce: 4 acc, fp32, stream st", 1, bufs, n, parts);
  run<4 | 8, 512>("2R+1W, no reduction, stream st, unroll 4", 4, bufs, n, parts);
  run<4, 1024>("2R+1W, no reduction, stream st", 2, bufs, n, parts);
  printf("done\n");
  return 0;
}
