// Developer harness (not part of the library): tcgen05 GEMM vs the fp32 CUDA-core GEMM on random data.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. -I../../../include dev/tc_gemm_test.cu -o tc_gemm_test
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tc_gemm_host.cuh"

namespace vldd {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
bool pdl_enabled() { return false; }
int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { set_error("%s: %s", what, cudaGetErrorString(e)); return VLDD_ERR_CUDA; }
  return 0;
}
}  // namespace vldd
using namespace vldd;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

static float* dev_rand(size_t n, unsigned seed, float scale = 1.0f) {
  std::vector<float> h(n);
  unsigned s = seed * 2654435761u + 12345u;
  for (size_t i = 0; i < n; ++i) { s = s * 1664525u + 1013904223u; h[i] = scale * (((s >> 8) & 0xFFFF) / 32768.0f - 1.0f); }
  float* d; CK(cudaMalloc(&d, n * 4)); CK(cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice));
  return d;
}

struct Res { double maxerr, maxref; float ms; };

template <bool AK, bool BKm, int kSplit, int kStagesT = 0, int BN = 128>
static Res run_case(const char* name, int M, int N, int K0, int K1, int splits, bool axpy) {
  const size_t asz0 = (size_t)M * K0, bsz0 = (size_t)N * K0, asz1 = (size_t)M * (K1 ? K1 : 1), bsz1 = (size_t)N * (K1 ? K1 : 1);
  float *A0 = dev_rand(asz0, 1), *B0 = dev_rand(bsz0, 2), *A1 = dev_rand(asz1, 3), *B1 = dev_rand(bsz1, 4);
  const int lda0 = AK ? K0 : M, ldb0 = BKm ? K0 : N, lda1 = AK ? K1 : M, ldb1 = BKm ? K1 : N;
  GemmOperands g = K1 ? gemm_ops2(A0, lda0, B0, ldb0, K0, A1, lda1, B1, ldb1, K1, M, N) : gemm_ops(A0, lda0, B0, ldb0, M, N, K0);
  const size_t csz = (size_t)M * N;
  float *Cref, *Ctc, *part, *src = dev_rand(csz, 5), *lr = dev_rand(4, 6);
  CK(cudaMalloc(&Cref, csz * 4)); CK(cudaMalloc(&Ctc, csz * 4)); CK(cudaMalloc(&part, csz * 4 * splits));
  CK(cudaMemset(Ctc, 0xFF, csz * 4));
  // reference
  if (axpy) launch_gemm<AK, BKm>(g, 1, nullptr, EpiAxpy{src, Cref, N, lr}, 0);
  else launch_gemm<AK, BKm>(g, 1, nullptr, EpiStore{Cref, N, 1.0f}, 0);
  CK(cudaDeviceSynchronize());
  if (!tc::gemm_ok<AK, BKm>(g)) { printf("%s: not eligible\n", name); exit(1); }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  static float* flush = nullptr;
  const size_t flush_bytes = 512ull << 20;
  if (getenv("COLD") && !flush) CK(cudaMalloc(&flush, flush_bytes));
  for (int rep = 0; rep < 5; ++rep) {
    if (flush) CK(cudaMemsetAsync(flush, rep, flush_bytes));
    CK(cudaEventRecord(e0));
    int rc;
    if (axpy) rc = tc::launch<AK, BKm, kSplit, tc::EpiAxpyTC, kStagesT, BN>(g, 1, tc::EpiAxpyTC{src, Ctc, N, lr}, 0);
    else if (splits > 1) rc = tc::launch<AK, BKm, kSplit, tc::EpiPartial, kStagesT>(g, splits, tc::EpiPartial{part, (long long)csz}, 0);
    else rc = tc::launch<AK, BKm, kSplit, tc::EpiScale, kStagesT, BN>(g, 1, tc::EpiScale{Ctc, N, 1.0f}, 0);
    if (rc) { printf("%s: launch failed: %s\n", name, g_err); exit(1); }
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  std::vector<float> href(csz), htc(csz), hp;
  CK(cudaMemcpy(href.data(), Cref, csz * 4, cudaMemcpyDeviceToHost));
  if (splits > 1 && !axpy) {
    hp.resize(csz * splits);
    CK(cudaMemcpy(hp.data(), part, csz * 4 * splits, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < csz; ++i) { float s = 0; for (int z = 0; z < splits; ++z) s += hp[z * csz + i]; htc[i] = s; }
  } else {
    CK(cudaMemcpy(htc.data(), Ctc, csz * 4, cudaMemcpyDeviceToHost));
  }
  double maxerr = 0, maxref = 0; size_t bad = 0;
  for (size_t i = 0; i < csz; ++i) {
    double e = fabs((double)href[i] - (double)htc[i]);
    if (!(e == e)) { e = 1e30; }
    if (e > maxerr) { maxerr = e; bad = i; }
    if (fabs(href[i]) > maxref) maxref = fabs(href[i]);
  }
  printf("%-34s M=%4d N=%4d K=%4d+%4d splits=%2d kSplit=%d  maxerr=%.3e maxref=%.3e rel=%.3e  (worst at m=%zu n=%zu ref=%f got=%f)  %.1f us\n",
         name, M, N, K0, K1, splits, kSplit, maxerr, maxref, maxerr / maxref, bad / N, bad % N, href[bad], htc[bad], ms * 1000);
  cudaFree(A0); cudaFree(B0); cudaFree(A1); cudaFree(B1); cudaFree(Cref); cudaFree(Ctc); cudaFree(part); cudaFree(src); cudaFree(lr);
  return {maxerr, maxref, ms};
}

int main(int argc, char** argv) {
  int only = argc > 1 ? atoi(argv[1]) : -1;
  int c = 0;
  auto want = [&](int id) { return only < 0 || only == id; };
  if (want(c++)) run_case<true, true, 1>("NT tiny 1xTF32", 128, 128, 32, 0, 1, false);
  if (want(c++)) run_case<true, true, 1>("NT small 1xTF32", 100, 256, 96, 0, 1, false);
  if (want(c++)) run_case<true, true, 3>("NT small 3xTF32", 100, 256, 96, 0, 1, false);
  if (want(c++)) run_case<true, true, 3>("NT F2 splitK", 100, 2304, 2304, 0, 8, false);
  if (want(c++)) run_case<true, true, 3>("NT F2t dual splitK", 100, 2304, 2304, 2304, 8, false);
  if (want(c++)) run_case<true, false, 1>("NN small 1x", 100, 256, 96, 0, 1, false);
  if (want(c++)) run_case<true, false, 3>("NN dh splitK", 100, 2304, 2304, 0, 8, false);
  if (want(c++)) run_case<false, false, 1>("TN small 1x", 256, 128, 100, 0, 1, false);
  if (want(c++)) run_case<false, false, 3>("TN dW2 axpy", 2304, 2304, 100, 0, 1, true);
  if (want(c++)) run_case<false, false, 3>("TN dW2t dual axpy", 2304, 2304, 100, 100, 1, true);
  if (want(c++)) run_case<false, false, 3>("TN dW1 axpy", 2304, 768, 100, 0, 1, true);
  if (want(c++)) run_case<true, true, 3>("NT sims 1000x5000x768", 1000, 5000, 768, 0, 1, false);
  if (want(c++)) run_case<true, true, 1>("NT sims 1x 1000x5000x768", 1000, 5000, 768, 0, 1, false);
  if (want(c++)) run_case<true, false, 3>("NN dY dual", 100, 768, 2304, 2304, 6, false);
  // stage-count / split-count experiments
  if (want(c++)) run_case<false, false, 3, 1>("TN dW2 axpy stages=1", 2304, 2304, 100, 0, 1, true);
  if (want(c++)) run_case<false, false, 3, 2>("TN dW2 axpy stages=2", 2304, 2304, 100, 0, 1, true);
  if (want(c++)) run_case<false, false, 3, 1>("TN dW2t dual axpy stages=1", 2304, 2304, 100, 100, 1, true);
  if (want(c++)) run_case<false, false, 3, 1>("TN dW1 axpy stages=1", 2304, 768, 100, 0, 1, true);
  if (want(c++)) run_case<true, true, 3, 1>("NT F2 splits=16 stages=1", 100, 2304, 2304, 0, 16, false);
  if (want(c++)) run_case<true, true, 3, 1>("NT F2 splits=24 stages=1", 100, 2304, 2304, 0, 24, false);
  if (want(c++)) run_case<true, true, 3, 2>("NT F2 splits=8 stages=2", 100, 2304, 2304, 0, 8, false);
  if (want(c++)) run_case<true, true, 3, 1>("NT F2 splits=8 stages=1", 100, 2304, 2304, 0, 8, false);
  if (want(c++)) run_case<true, true, 3, 1>("NT sims stages=1", 1000, 5000, 768, 0, 1, false);
  if (want(c++)) run_case<true, true, 1, 2>("NT F2 1xTF32 splits=8 stages=2", 100, 2304, 2304, 0, 8, false);
  if (want(c++)) run_case<true, true, 1, 2>("NT F2 1xTF32 splits=16 stages=2", 100, 2304, 2304, 0, 16, false);
  if (want(c++)) run_case<false, false, 3, 0, 96>("TN dW2 axpy BN=96", 2304, 2304, 100, 0, 1, true);
  if (want(c++)) run_case<false, false, 3, 0, 96>("TN dW2t dual axpy BN=96", 2304, 2304, 100, 100, 1, true);
  if (want(c++)) run_case<true, true, 3, 0, 96>("NT sims BN=96 1000x4992x768", 1000, 4992, 768, 0, 1, false);
  if (want(c++)) run_case<true, false, 3, 0, 64>("NN small BN=64", 100, 256, 96, 0, 1, false);
  if (want(c++)) run_case<true, true, 1>("NT big 1xTF32 8192x8192x2048", 8192, 8192, 2048, 0, 1, false);
  if (want(c++)) run_case<true, true, 3>("NT big 3xTF32 8192x8192x2048", 8192, 8192, 2048, 0, 1, false);
  printf("done\n");
  return 0;
}
