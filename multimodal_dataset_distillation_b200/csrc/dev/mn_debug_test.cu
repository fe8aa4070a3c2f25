// Developer harness: probe the MN-major descriptor semantics (LBO / SBO / k-step) of tcgen05.mma kind::tf32.
#define VLDD_TC_DEBUG 1
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
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

int main() {
  // B MN-major: C[m,n] = sum_k A[m,k] B[k,n]; A = one-hot rows so C[m,:] = B[k(m),:] -> reveals which B element lands where
  const int M = 128, N = 128, K = 32;
  std::vector<float> hA((size_t)M * K, 0.f), hB((size_t)K * N);
  for (int m = 0; m < M; ++m) hA[(size_t)m * K + (m % K)] = 1.0f;          // row m picks k = m % 32
  for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) hB[(size_t)k * N + n] = k * 1000 + n;
  float *A, *B, *C;
  CK(cudaMalloc(&A, hA.size() * 4)); CK(cudaMalloc(&B, hB.size() * 4)); CK(cudaMalloc(&C, (size_t)M * N * 4));
  CK(cudaMemcpy(A, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice));
  GemmOperands g = gemm_ops(A, K, B, N, M, N, K);
  const uint32_t trials[][3] = {{4096, 512, 1024}, {512, 4096, 1024}, {4096, 1024, 1024}};
  for (auto& t : trials) {
    uint32_t dbg[4] = {t[0], t[1], t[2], 0};
    CK(cudaMemcpyToSymbol(tc::g_dbg, dbg, sizeof dbg));
    CK(cudaMemset(C, 0, (size_t)M * N * 4));
    if (tc::launch<true, false, 1>(g, 1, tc::EpiScale{C, N, 1.0f}, 0)) { printf("launch failed %s\n", g_err); return 1; }
    CK(cudaDeviceSynchronize());
    std::vector<float> h((size_t)M * N);
    CK(cudaMemcpy(h.data(), C, h.size() * 4, cudaMemcpyDeviceToHost));
    int good = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) good += (h[(size_t)m * N + n] == (m % K) * 1000 + n);
    printf("lbo=%u sbo=%u step=%u: %d / %d correct\n", t[0], t[1], t[2], good, M * N);
    for (int m : {0, 1, 8, 9, 31}) {
      printf("   m=%2d:", m);
      for (int n : {0, 1, 4, 31, 32, 33, 64, 96, 127}) printf(" [%d]=%.0f", n, h[(size_t)m * N + n]);
      printf("\n");
    }
  }
  return 0;
}
