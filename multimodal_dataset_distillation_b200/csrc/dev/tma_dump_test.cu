// Developer harness: what does the 3-D (MN-major) tensor map put into shared memory?
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tc_gemm_host.cuh"
namespace vldd {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
bool pdl_enabled() { return false; }
int check_launch(const char*) { return 0; }
}
using namespace vldd;
using namespace vldd::tc;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__global__ void dump_kernel(const __grid_constant__ CUtensorMap map, int is3d, int c0, int c1, int c2, float* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  for (int i = threadIdx.x; i < TILE_BYTES / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = -777.0f;
  __syncthreads();
  if (threadIdx.x == 0) {
    fence_proxy_async();
    mbar_expect_tx(&bar, TILE_BYTES);
    if (is3d) tma_load_3d(smem, &map, &bar, c0, c1, c2); else tma_load_2d(smem, &map, &bar, c0, c1);
  }
  mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < TILE_BYTES / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}

int main() {
  const int K = 96, MN = 256;
  std::vector<float> h((size_t)K * MN);
  for (int k = 0; k < K; ++k) for (int n = 0; n < MN; ++n) h[(size_t)k * MN + n] = k * 1000 + n;   // value encodes (k, n)
  float* d; CK(cudaMalloc(&d, h.size() * 4)); CK(cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  float* out; CK(cudaMalloc(&out, TILE_BYTES));
  CUtensorMap m;
  if (get_map(d, MN, K, MN, false, &m)) { printf("map failed %s\n", g_err); return 1; }
  CK(cudaFuncSetAttribute(dump_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_BYTES + 1024));
  for (int trial = 0; trial < 2; ++trial) {
    const int k0 = trial ? 32 : 0, chunk0 = trial ? 4 : 0;
    dump_kernel<<<1, 128, TILE_BYTES + 1024>>>(m, 1, 0, k0, chunk0, out);
    CK(cudaDeviceSynchronize());
    std::vector<float> o(TILE_BYTES / 4);
    CK(cudaMemcpy(o.data(), out, TILE_BYTES, cudaMemcpyDeviceToHost));
    int bad = 0;
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 32; ++k) for (int i = 0; i < 32; ++i) {
      size_t byte = ((size_t)c * 1024 + k * 32 + i) * 4;
      size_t sw = byte ^ (((byte >> 7) & 7) << 4);
      float expect = (k0 + k) * 1000 + (chunk0 + c) * 32 + i;
      if (o[sw / 4] != expect) { if (bad < 8) printf("trial %d mismatch c=%d k=%d i=%d expect %.0f got %.0f\n", trial, c, k, i, expect, o[sw / 4]); ++bad; }
    }
    printf("trial %d: %d mismatches; first floats: %.0f %.0f %.0f %.0f | row1: %.0f %.0f | chunk1: %.0f\n", trial, bad, o[0], o[1], o[2], o[3], o[32], o[36], o[1024]);
  }
  return 0;
}
