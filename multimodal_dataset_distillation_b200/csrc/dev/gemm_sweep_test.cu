// Developer harness (not part of the library): tile width / pipeline depth / split-K sweep of the tcgen05 GEMM on the
// shapes of one distill step.  Every case is timed the way the engine runs it: a dependent chain
//   GEMM (split-K partial slabs) -> consumer that sums the slabs -> GEMM -> ...
// with programmatic dependent launch, L2-warm operands, CUDA events around the whole chain.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. -I../../../include dev/gemm_sweep_test.cu -o dev/gemm_sweep_test
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../tc_gemm_host.cuh"

namespace vldd {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
bool pdl_enabled() { return true; }
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

// stand-in for the engine's epi_* / row kernels: out[i] = sum_z part[z][i], all slab loads in flight before the adds
__global__ void __launch_bounds__(256) consume_kernel(const float* __restrict__ part, int splits, size_t stride, size_t n4,
                                                      float* __restrict__ out) {
  pdl_enter();
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < splits; ++z) {
      const float4 v = reinterpret_cast<const float4*>(part + (size_t)z * stride)[i];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    reinterpret_cast<float4*>(out)[i] = acc;
  }
}

struct Shape { const char* name; int M, N, K0, K1; };

template <bool AK, bool BKm, int BN, int ST, int EW = 4>
static void run_partial(const Shape& sh, int splits, const float* ref, float* refbuf_host) {
  const int M = sh.M, N = sh.N, K0 = sh.K0, K1 = sh.K1;
  const size_t asz0 = (size_t)M * K0, bsz0 = (size_t)N * K0, asz1 = (size_t)M * (K1 ? K1 : 1), bsz1 = (size_t)N * (K1 ? K1 : 1);
  float *A0 = dev_rand(asz0, 1), *B0 = dev_rand(bsz0, 2), *A1 = dev_rand(asz1, 3), *B1 = dev_rand(bsz1, 4);
  const int lda0 = AK ? K0 : M, ldb0 = BKm ? K0 : N, lda1 = AK ? K1 : M, ldb1 = BKm ? K1 : N;
  GemmOperands g = K1 ? gemm_ops2(A0, lda0, B0, ldb0, K0, A1, lda1, B1, ldb1, K1, M, N) : gemm_ops(A0, lda0, B0, ldb0, M, N, K0);
  const size_t csz = (size_t)M * N;
  if (splits <= 0) splits = tc::pick_splits(M, N, K0 + K1, BN);
  float *part, *out, *Cref;
  CK(cudaMalloc(&part, csz * 4 * splits)); CK(cudaMalloc(&out, csz * 4)); CK(cudaMalloc(&Cref, csz * 4));
  launch_gemm<AK, BKm>(g, 1, nullptr, EpiStore{Cref, N, 1.0f}, 0);
  CK(cudaDeviceSynchronize());
  auto one = [&]() {
    int rc = tc::launch<AK, BKm, 3, tc::EpiPartial, ST, BN, EW>(g, splits, tc::EpiPartial{part, (long long)csz}, 0);
    if (rc) { printf("launch failed: %s\n", g_err); exit(1); }
    launch_k(consume_kernel, 148, 256, 0, (cudaStream_t)0, (const float*)part, splits, csz, csz / 4, out);
  };
  for (int i = 0; i < 5; ++i) one();
  CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  std::vector<float> hr(csz), ho(csz);
  CK(cudaMemcpy(hr.data(), Cref, csz * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ho.data(), out, csz * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < csz; ++i) {
    double e = fabs((double)hr[i] - (double)ho[i]); if (!(e == e)) e = 1e30;
    if (e > maxerr) maxerr = e;
    if (fabs(hr[i]) > maxref) maxref = fabs(hr[i]);
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int R = 40;
  float best = 1e30f, sum = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < R; ++i) one();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best; sum += ms;
  }
  // GEMM alone, back to back
  float best_g = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < R; ++i) tc::launch<AK, BKm, 3, tc::EpiPartial, ST, BN, EW>(g, splits, tc::EpiPartial{part, (long long)csz}, 0);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best_g = ms < best_g ? ms : best_g;
  }
  using C = tc::Cfg<3, true, ST, BN, EW>;
  printf("%-10s M=%4d N=%4d K=%4d+%4d  BN=%3d stages=%d ew=%d splits=%2d ctas=%3d kb/cta=%4.1f | pair %6.2f us (avg %6.2f)  gemm %6.2f us | rel err %.1e\n",
         sh.name, M, N, K0, K1, BN, C::kStages, EW, splits, ((M + 127) / 128) * ((N + BN - 1) / BN) * splits,
         (double)((K0 + 31) / 32 + (K1 + 31) / 32) / splits, best * 1000 / R, sum * 1000 / R / 5, best_g * 1000 / R, maxerr / maxref);
  fflush(stdout);
  cudaFree(A0); cudaFree(B0); cudaFree(A1); cudaFree(B1); cudaFree(part); cudaFree(out); cudaFree(Cref);
}

template <int BN, int ST, int EW = 4>
static void run_axpy(const Shape& sh) {
  const int M = sh.M, N = sh.N, K0 = sh.K0, K1 = sh.K1;
  // TN: A element (m,k) at A[k*M + m], B element (k,n) at B[k*N + n]
  float *A0 = dev_rand((size_t)M * K0, 1), *B0 = dev_rand((size_t)N * K0, 2), *A1 = dev_rand((size_t)M * (K1 ? K1 : 1), 3),
        *B1 = dev_rand((size_t)N * (K1 ? K1 : 1), 4);
  GemmOperands g = K1 ? gemm_ops2(A0, M, B0, N, K0, A1, M, B1, N, K1, M, N) : gemm_ops(A0, M, B0, N, M, N, K0);
  const size_t csz = (size_t)M * N;
  float *src = dev_rand(csz, 5), *lr = dev_rand(4, 6), *dst, *Cref;
  CK(cudaMalloc(&dst, csz * 4)); CK(cudaMalloc(&Cref, csz * 4));
  launch_gemm<false, false>(g, 1, nullptr, EpiAxpy{src, Cref, N, lr}, 0);
  CK(cudaDeviceSynchronize());
  auto one = [&]() {
    int rc = tc::launch<false, false, 3, tc::EpiAxpyTC, ST, BN, EW>(g, 1, tc::EpiAxpyTC{src, dst, N, lr}, 0);
    if (rc) { printf("launch failed: %s\n", g_err); exit(1); }
  };
  for (int i = 0; i < 3; ++i) one();
  CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  std::vector<float> hr(csz), ho(csz);
  CK(cudaMemcpy(hr.data(), Cref, csz * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ho.data(), dst, csz * 4, cudaMemcpyDeviceToHost));
  double maxerr = 0, maxref = 0;
  for (size_t i = 0; i < csz; ++i) {
    double e = fabs((double)hr[i] - (double)ho[i]); if (!(e == e)) e = 1e30;
    if (e > maxerr) maxerr = e;
    if (fabs(hr[i]) > maxref) maxref = fabs(hr[i]);
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int R = 20;
  float best = 1e30f;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    for (int i = 0; i < R; ++i) one();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  using C = tc::Cfg<3, true, ST, BN, EW>;
  printf("%-10s M=%4d N=%4d K=%4d+%4d  BN=%3d stages=%d ew=%d axpy tiles=%3d | %6.2f us | rel err %.1e\n", sh.name, M, N, K0, K1, BN,
         C::kStages, EW, ((M + 127) / 128) * ((N + BN - 1) / BN), best * 1000 / R, maxerr / maxref);
  fflush(stdout);
  cudaFree(A0); cudaFree(B0); cudaFree(A1); cudaFree(B1); cudaFree(src); cudaFree(lr); cudaFree(dst); cudaFree(Cref);
}

template <int kSplit, int BN, int ST, int EW>
static void run_big(int M, int N, int K) {
  float *A = dev_rand((size_t)M * K, 1), *B = dev_rand((size_t)N * K, 2), *Cc;
  CK(cudaMalloc(&Cc, (size_t)M * N * 4));
  GemmOperands g = gemm_ops(A, K, B, K, M, N, K);
  auto one = [&]() {
    int rc = tc::launch<true, true, kSplit, tc::EpiScale, ST, BN, EW>(g, 1, tc::EpiScale{Cc, N, 1.0f}, 0);
    if (rc) { printf("launch failed: %s\n", g_err); exit(1); }
  };
  one(); one();
  CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0));
    one();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  using C = tc::Cfg<kSplit, kSplit == 3, ST, BN, EW>;
  printf("big %dx%dx%d kSplit=%d BN=%d stages=%d ew=%d | %8.1f us  %6.1f TFLOP/s useful, %6.1f TFLOP/s of MMA work\n", M, N, K, kSplit, BN,
         C::kStages, EW, best * 1000, 2.0 * M * N * K / (best * 1e-3) / 1e12, kSplit * 2.0 * M * N * K / (best * 1e-3) / 1e12);
  fflush(stdout);
  cudaFree(A); cudaFree(B); cudaFree(Cc);
}

// MMA rate by operand layout (timing only): K-major operands are [rows][K], MN-major ones [K][rows]
template <bool AK, bool BKm>
static void run_big_layout(int M, int N, int K) {
  float *A = dev_rand((size_t)M * K, 1), *B = dev_rand((size_t)N * K, 2), *Cc;
  CK(cudaMalloc(&Cc, (size_t)M * N * 4));
  GemmOperands g = gemm_ops(A, AK ? K : M, B, BKm ? K : N, M, N, K);
  auto one = [&]() {
    int rc = tc::launch<AK, BKm, 3, tc::EpiScale, 0, 128, 4>(g, 1, tc::EpiScale{Cc, N, 1.0f}, 0);
    if (rc) { printf("launch failed: %s\n", g_err); exit(1); }
  };
  one(); one();
  CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0)); one(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  printf("big %dx%dx%d 3xTF32 A %s-major, B %s-major | %8.1f us  %6.1f TFLOP/s of MMA work\n", M, N, K, AK ? "K" : "MN", BKm ? "K" : "MN",
         best * 1000, 3 * 2.0 * M * N * K / (best * 1e-3) / 1e12);
  fflush(stdout);
  cudaFree(A); cudaFree(B); cudaFree(Cc);
}

template <int BN>
static void run_big_bf16(int M, int N, int K) {
  uint16_t *Ah, *Al, *Bh, *Bl; float* Cc;
  CK(cudaMalloc(&Ah, (size_t)M * K * 2)); CK(cudaMalloc(&Al, (size_t)M * K * 2)); CK(cudaMalloc(&Bh, (size_t)N * K * 2)); CK(cudaMalloc(&Bl, (size_t)N * K * 2));
  CK(cudaMemset(Ah, 0x3c, (size_t)M * K * 2)); CK(cudaMemset(Al, 0x38, (size_t)M * K * 2)); CK(cudaMemset(Bh, 0x3c, (size_t)N * K * 2)); CK(cudaMemset(Bl, 0x38, (size_t)N * K * 2));
  CK(cudaMalloc(&Cc, (size_t)M * N * 4));
  auto one = [&]() {
    int rc = tc::launch_bf16x3<tc::EpiScale, BN>(Ah, Al, Bh, Bl, M, N, K, tc::EpiScale{Cc, N, 1.0f}, 0);
    if (rc) { printf("launch failed: %s\n", g_err); exit(1); }
  };
  one(); one();
  CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    CK(cudaEventRecord(e0)); one(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    best = ms < best ? ms : best;
  }
  printf("big %dx%dx%d bf16x3 BN=%d full store epilogue | %8.1f us  %6.1f TFLOP/s useful, %6.1f TFLOP/s of MMA work\n", M, N, K, BN, best * 1000,
         2.0 * M * N * K / (best * 1e-3) / 1e12, 3 * 2.0 * M * N * K / (best * 1e-3) / 1e12);
  fflush(stdout);
  cudaFree(Ah); cudaFree(Al); cudaFree(Bh); cudaFree(Bl); cudaFree(Cc);
}

template <bool AK, bool BKm>
static void sweep_partial(const Shape& sh) {
  run_partial<AK, BKm, 128, 0>(sh, 0, nullptr, nullptr);
  run_partial<AK, BKm, 64, 0>(sh, 0, nullptr, nullptr);
  run_partial<AK, BKm, 128, 0, 8>(sh, 0, nullptr, nullptr);
  run_partial<AK, BKm, 64, 0, 8>(sh, 0, nullptr, nullptr);
  run_partial<AK, BKm, 96, 0, 8>(sh, 0, nullptr, nullptr);
}

int main(int argc, char** argv) {
  if (argc > 1 && argv[1][0] == 'a') {       // "axpy": the weight-gradient GEMMs only (build with -DVLDD_EXP_NOSRC / _NOSTORE to take the epilogue apart)
    const int B = 100, dt = 768, d = 2304;
    const Shape dW2{"dW2", d, d, B, 0}, dW2k32{"dW2k32", d, d, 32, 0}, dW2t{"dW2t", d, d, B, B}, dW1{"dW1", d, dt, B, 0}, dW1k32{"dW1k32", d, dt, 32, 0};
    run_axpy<96, 0, 8>(dW2); run_axpy<96, 0, 8>(dW2k32); run_axpy<96, 0, 4>(dW2t); run_axpy<96, 0, 8>(dW2t); run_axpy<96, 0, 8>(dW1); run_axpy<96, 0, 8>(dW1k32);
    printf("done\n");
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'f') {       // "flickr": the 1000 x 5000 x 768 score GEMM by tile width
    run_big<3, 128, 0, 4>(1000, 5000, 768); run_big<3, 96, 0, 4>(1000, 5000, 768); run_big<3, 64, 0, 4>(1000, 5000, 768);
    run_big<3, 128, 0, 8>(1000, 5000, 768); run_big<3, 96, 0, 8>(1000, 5000, 768); run_big<3, 64, 0, 8>(1000, 5000, 768);
    run_big<3, 96, 0, 4>(1000, 5000, 2304); run_big<3, 128, 0, 4>(1000, 5000, 2304);
    printf("done\n");
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'l') {       // "layout": tensor rate of the four operand layouts
    run_big_layout<true, true>(8192, 8192, 2048); run_big_layout<true, false>(8192, 8192, 2048);
    run_big_layout<false, false>(8192, 8192, 2048); run_big_layout<false, true>(8192, 8192, 2048);
    printf("done\n");
    return 0;
  }
  if (argc > 1 && argv[1][0] == 'b') {       // "big": retrieval-shaped products only
    run_big<3, 128, 0, 4>(5000, 25000, 768);
    run_big<1, 128, 0, 4>(5000, 25000, 768);
    run_big_bf16<128>(5000, 25000, 768);
    run_big_bf16<256>(5000, 25000, 768);
    run_big<3, 128, 0, 4>(8192, 8192, 2048);
    run_big_bf16<128>(8192, 8192, 2048);
    run_big_bf16<256>(8192, 8192, 2048);
    printf("done\n");
    return 0;
  }
  const int B = 100, dt = 768, d = 2304;
  const Shape p{"p", B, d, dt, 0}, f{"f", B, d, d, 0}, fd{"fd", B, d, d, d}, S{"S", B, B, d, 0}, dh{"dh", B, d, d, 0},
      dhd{"dhd", B, d, d, d}, dY{"dY", B, dt, d, d};
  printf("== K-major x K-major (p, f, fd, S)\n");
  sweep_partial<true, true>(p);
  sweep_partial<true, true>(f);
  sweep_partial<true, true>(fd);
  sweep_partial<true, true>(S);
  printf("== K-major x MN-major (dh, dhd, dY)\n");
  sweep_partial<true, false>(dh);
  sweep_partial<true, false>(dhd);
  sweep_partial<true, false>(dY);
  printf("== weight-gradient GEMMs (MN x MN, axpy epilogue)\n");
  const Shape dW2{"dW2", d, d, B, 0}, dW2t{"dW2t", d, d, B, B}, dW1{"dW1", d, dt, B, 0};
  run_axpy<96, 5>(dW2); run_axpy<96, 0, 8>(dW2); run_axpy<96, 3, 8>(dW2); run_axpy<128, 0, 8>(dW2); run_axpy<64, 0, 8>(dW2);
  run_axpy<96, 5>(dW2t); run_axpy<96, 0, 8>(dW2t); run_axpy<96, 3, 8>(dW2t); run_axpy<128, 0, 8>(dW2t); run_axpy<64, 0, 8>(dW2t);
  run_axpy<96, 5>(dW1); run_axpy<96, 0, 8>(dW1); run_axpy<128, 0, 8>(dW1); run_axpy<64, 0, 8>(dW1);
  printf("== big K-major GEMM, full store epilogue (retrieval-like)\n");
  run_big<3, 128, 0, 4>(8192, 8192, 2048); run_big<3, 128, 0, 8>(8192, 8192, 2048);
  run_big<1, 128, 0, 4>(8192, 8192, 2048); run_big<1, 128, 0, 8>(8192, 8192, 2048);
  run_big<3, 128, 0, 4>(5000, 25000, 768); run_big<3, 128, 0, 8>(5000, 25000, 768);
  run_big<1, 128, 0, 4>(5000, 25000, 768); run_big<1, 128, 0, 8>(5000, 25000, 768);
  printf("done\n");
  return 0;
}
