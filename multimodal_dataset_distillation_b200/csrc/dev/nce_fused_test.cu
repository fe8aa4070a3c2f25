// Developer harness (not part of the library): the CUDA-core InfoNCE kernels (nce_fused.cuh) alone, Flickr shape, with a
// globaltimer stamp at every phase boundary of one CTA and CUDA-event times of a dependent chain of launches.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -DVLDD_NCE_TIMELINE -I. -I../../include dev/nce_fused_test.cu -o dev/nce_fused_test
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../nce_fused.cuh"

namespace vldd {
static char g_err[512];
void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
bool pdl_enabled() { return true; }
int check_launch(const char*) { return 0; }
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

int main() {
  const int B = 100, d = 2304, ld = 128;
  float *X = dev_rand((size_t)B * d, 1, 0.02f), *Y = dev_rand((size_t)B * d, 2, 0.02f), *scale = dev_rand(4, 3), *lr = dev_rand(4, 4);
  float *S, *Sd, *G, *Pr, *Pc, *Gd, *lse_r, *lse_c, *loss, *out, *dl;
  CK(cudaMalloc(&S, B * ld * 4)); CK(cudaMalloc(&Sd, B * ld * 4)); CK(cudaMalloc(&G, B * ld * 4)); CK(cudaMalloc(&Pr, B * ld * 4));
  CK(cudaMalloc(&Pc, B * ld * 4)); CK(cudaMalloc(&Gd, B * ld * 4)); CK(cudaMalloc(&lse_r, 512)); CK(cudaMalloc(&lse_c, 512));
  CK(cudaMalloc(&loss, 16)); CK(cudaMalloc(&out, (size_t)B * d * 4)); CK(cudaMalloc(&dl, 16));
  CK(cudaMemset(S, 0, B * ld * 4)); CK(cudaMemset(Sd, 0, B * ld * 4)); CK(cudaMemset(dl, 0, 16));
  const int cap = 224 * 1024;
  CK(cudaFuncSetAttribute(small_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  CK(cudaFuncSetAttribute(nce_gx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  CK(cudaFuncSetAttribute(nce_t_gx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, cap));
  auto scores = [&](float* dst) {
    launch_k(small_scores_kernel, dim3(ceil_div(B, kScoresTJ), ceil_div(B, kScoresTI)), kScoresThreads, scores_smem_bytes(d), (cudaStream_t)0,
             (const float*)X, (const float*)Y, B, d, (const float*)scale, dst, ld);
  };
  auto gx = [&]() {
    launch_k(nce_gx_kernel, ceil_div(d, kNceGxCols), kNceGxThreads, nce_gx_smem_bytes(B, ld), (cudaStream_t)0, (const float*)S, (const float*)X, B,
             ld, d, lse_r, lse_c, G, Pr, Pc, loss, out);
  };
  auto tgx = [&]() {
    launch_k(nce_t_gx_kernel, ceil_div(d, kNceGxCols), kNceGxThreads, nce_gx_smem_bytes(B, ld), (cudaStream_t)0, (const float*)S, (const float*)Sd,
             (const float*)Pr, (const float*)Pc, (const float*)X, B, ld, d, (const float*)lr, (const float*)scale, Gd, dl, dl + 1, out);
  };
  scores(S); scores(Sd); gx(); tgx();
  CK(cudaDeviceSynchronize()); CK(cudaGetLastError());
  long long tl[3][16];
  CK(cudaMemcpyFromSymbol(tl, g_nce_tl, sizeof tl));
  const char* names[3] = {"small_scores", "nce_gx", "nce_t_gx"};
  for (int k = 0; k < 3; ++k) {
    printf("%-14s phases (ns since kernel entry of CTA 5):", names[k]);
    for (int s = 1; s <= 6; ++s) if (tl[k][s]) printf(" %6lld", tl[k][s] - tl[k][0]);
    printf("\n");
  }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto time_it = [&](const char* what, auto fn) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      CK(cudaEventRecord(e0));
      for (int i = 0; i < 20; ++i) fn();
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      best = ms < best ? ms : best;
    }
    printf("%-28s %7.2f us per launch\n", what, best * 1000 / 20);
  };
  time_it("small_scores", [&]() { scores(S); });
  time_it("nce_gx", gx);
  time_it("nce_t_gx", tgx);
  time_it("scores + nce_gx (pair)", [&]() { scores(S); gx(); });
  time_it("scores + nce_t_gx (pair)", [&]() { scores(Sd); tgx(); });
  printf("done\n");
  return 0;
}
