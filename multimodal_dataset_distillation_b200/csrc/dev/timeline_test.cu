// Developer harness: phase timeline (globaltimer, ns) of the skinny split-K GEMM and the weight-gradient GEMM.
#define VLDD_TC_TIMELINE 1
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tc_gemm_host.cuh"
namespace vldd {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
int check_launch(const char*) { return 0; }
bool pdl_enabled() { return false; }
}
using namespace vldd;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)
static float* dev_rand(size_t n) { std::vector<float> h(n); for (size_t i = 0; i < n; ++i) h[i] = (float)((i * 2654435761u) % 1000) / 1000.f - 0.5f; float* d; CK(cudaMalloc(&d, n * 4)); CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice)); return d; }

static void report(const char* name, int nctas) {
  std::vector<long long> tl(148 * 10 * 16);
  CK(cudaMemcpyFromSymbol(tl.data(), tc::g_timeline, tl.size() * 8));
  long long t0 = (1ll << 62);
  for (int c = 0; c < nctas; ++c) if (tl[(c * 10 + 0) * 16 + 0] && tl[(c * 10 + 0) * 16 + 0] < t0) t0 = tl[(c * 10 + 0) * 16 + 0];
  printf("%s (ns since first CTA start; cta: start prologue_done | producer first_issue all_issued | mma first_ready last_commit | splitter first_full first_done all_done | epi tmem_full done | exit)\n", name);
  for (int c : {0, 1, nctas / 2, nctas - 1}) {
    auto T = [&](int w, int s) { long long v = tl[(c * 10 + w) * 16 + s]; return v ? (long long)(v - t0) : -1ll; };
    printf("  cta %3d: %6lld %6lld | %6lld %6lld | %6lld %6lld | %6lld %6lld %6lld | %6lld %6lld | %6lld\n", c, T(0, 0), T(0, 2), T(0, 3), T(0, 4),
           T(1, 3), T(1, 4), T(2, 3), T(2, 4), T(2, 5), T(6, 6), T(6, 7), T(0, 8));
  }
}

int main() {
  const int M = 100, N = 2304, K = 2304;
  float *A = dev_rand((size_t)M * K), *W = dev_rand((size_t)N * K), *part;
  CK(cudaMalloc(&part, (size_t)8 * M * N * 4));
  GemmOperands g = gemm_ops(A, K, W, K, M, N, K);
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemset(part, 0, 16));
    if (tc::launch<true, true, 3>(g, 8, tc::EpiPartial{part, (long long)M * N}, 0)) { printf("fail %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
  }
  report("F2 skinny split-K 8 (144 CTAs)", 144);
  // weight-gradient GEMM
  const int Md = 2304, Nd = 2304, Kd = 100;
  float *dA = dev_rand((size_t)Kd * Md), *dB = dev_rand((size_t)Kd * Nd), *src = dev_rand((size_t)Md * Nd), *dst, *lr = dev_rand(4);
  CK(cudaMalloc(&dst, (size_t)Md * Nd * 4));
  GemmOperands g2 = gemm_ops(dA, Md, dB, Nd, Md, Nd, Kd);
  for (int rep = 0; rep < 3; ++rep) {
    if (tc::launch<false, false, 3>(g2, 1, tc::EpiAxpyTC{src, dst, Nd, lr}, 0)) { printf("fail %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
  }
  report("dW2 axpy (324 tiles on 148 persistent CTAs; epilogue columns = first tile tmem_full / all tiles done)", 148);
  // the engine's configuration of the same product: 96-wide tiles, 8 epilogue warps; per-tile stamps
  for (int rep = 0; rep < 3; ++rep) {
    if (tc::launch<false, false, 3, tc::EpiAxpyTC, 0, 96, 8>(g2, 1, tc::EpiAxpyTC{src, dst, Nd, lr}, 0)) { printf("fail %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
  }
  {
    std::vector<long long> tl(148 * 10 * 16);
    CK(cudaMemcpyFromSymbol(tl.data(), tc::g_timeline, tl.size() * 8));
    long long t0 = (1ll << 62);
    for (int c = 0; c < 148; ++c) if (tl[(c * 10 + 0) * 16 + 0] && tl[(c * 10 + 0) * 16 + 0] < t0) t0 = tl[(c * 10 + 0) * 16 + 0];
    printf("dW2 axpy, BN=96, 8 epilogue warps, 432 tiles (ns since first CTA start)\n");
    printf("  cta: start | mma: accumulator of tile 0,1,2 committed | epilogue warp 6 (chunks 0,64): tile0 begin end, tile1 begin end, tile2 begin end | warp 9 | exit\n");
    for (int c : {0, 1, 74, 147}) {
      auto T = [&](int w, int s) { long long v = tl[(c * 10 + w) * 16 + s]; return v ? (long long)(v - t0) : -1ll; };
      printf("  cta %3d: %6lld | %6lld %6lld %6lld | %6lld %6lld  %6lld %6lld  %6lld %6lld | %6lld %6lld  %6lld %6lld  %6lld %6lld | %6lld\n", c, T(0, 0),
             T(1, 8), T(1, 10), T(1, 12), T(6, 8), T(6, 9), T(6, 10), T(6, 11), T(6, 12), T(6, 13), T(9, 8), T(9, 9), T(9, 10), T(9, 11), T(9, 12), T(9, 13), T(0, 8));
    }
  }
  return 0;
}
