// Internal declarations of the unroll engine (public C ABI: include/vldd_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace vldd {

size_t unrolled_match_workspace_bytes(int N, int B, int K, int dt, int d);
int unrolled_match(const float* theta0, const float* theta_tgt, const float* Y, const float* U, const float* lr,
                   const float* scale, const int64_t* perms, const float* masks, float dropout_p,
                   unsigned long long* rng_state, int N, int B, int K, int dt, int d, float* out5, float* ce, float* dY,
                   float* dU, float* theta_K, void* workspace, size_t workspace_bytes, cudaStream_t st);
size_t contrastive_step_workspace_bytes(int B, int dt, int d);
int contrastive_step(const float* theta, const float* Y, const float* U, const float* scale, const float* mask, int B,
                     int dt, int d, float* loss, float* g_theta, float* dY, float* dU, float* dscale, void* workspace,
                     size_t workspace_bytes, cudaStream_t st);
int clip_loss(const float* theta, const float* Y, const float* U, const float* scale, const float* mask, int B, int dt,
              int d, float* loss, int32_t* top1, float* g_theta, float* dY, float* dU, float* dscale, void* workspace,
              size_t workspace_bytes, cudaStream_t st);
size_t infonce_workspace_bytes(int B, int d);
int infonce_grad(const float* xn, const float* yn, const float* scale, int B, int d, float* loss, float* dxn, float* dyn,
                 float* dscale, void* workspace, size_t workspace_bytes, cudaStream_t st);
int infonce_hvp(const float* xn, const float* yn, const float* scale, const float* cx, const float* cy, const float* cs,
                int B, int d, float* Ldot, float* hx, float* hy, float* hs, void* workspace, size_t workspace_bytes,
                cudaStream_t st);
size_t proj_head_workspace_bytes(int rows, int dt, int d);
int proj_head_forward(const float* theta, const float* Y, const float* mask, int rows, int dt, int d, float* z,
                      float* zn, void* workspace, size_t workspace_bytes, cudaStream_t st);

size_t skinny_gemm_workspace_bytes(int M, int N, int K);
int skinny_gemm_partial(const float* A, const float* W, int M, int N, int K, float* partial, size_t partial_bytes,
                        int* splits, cudaStream_t st);

}  // namespace vldd
