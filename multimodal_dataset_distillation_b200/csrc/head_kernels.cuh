// Row-wise / element-wise kernels of the text_projection head and the bidirectional InfoNCE loss, in primal
// and tangent (forward-over-reverse) form.  Every formula is the one in oracle/distill_ref.py
// (step_first_order / step_tangent), which is checked against torch double-backward in float64.
//
//   reference sites: networks.py:639-646 (ProjectionHead.forward), distill.py:546 (row normalise),
//   distill.py:548-551 (logits + 2x cross_entropy), distill.py:562-567 (autograd.grad create_graph=True).
//
// GEMM results arrive as split-K partial slabs [splits][rows*cols]; the consumer sums the slabs in fixed
// order (deterministic) and applies the fused math.
#pragma once
#include "common.cuh"

namespace vldd {

constexpr float kLnEps = 1e-5f;  // nn.LayerNorm default (networks.py:637)

// float4 helpers (shared with row_kernels_v4.cuh)
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float a) { return make_float4(a, a, a, a); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }

// Slab sums.  These kernels are latency-bound (a few hundred KB per launch, L2-resident), so what matters is how many
// L2 round trips sit on the critical path: every variant issues ALL its slab loads before the first addition
// (predicated, fully unrolled up to kMaxSlabsMlp) and then adds in the fixed order s = 0, 1, 2, ... -- the same
// association as a plain sequential loop, so results do not depend on which variant ran.
constexpr int kMaxSlabsMlp = 8;      // tc::pick_splits never exceeds 8 for the skinny weight GEMMs (18 tiles on 148 SMs)

__device__ __forceinline__ float sum_slabs(const float* __restrict__ part, int splits, size_t stride, size_t idx) {
  if (splits <= kMaxSlabsMlp) {
    float t[kMaxSlabsMlp];
#pragma unroll
    for (int s = 0; s < kMaxSlabsMlp; ++s) t[s] = s < splits ? part[(size_t)s * stride + idx] : 0.f;
    float v = t[0];
#pragma unroll
    for (int s = 1; s < kMaxSlabsMlp; ++s) v += t[s];
    return v;
  }
  float v = part[idx];
  for (int s = 1; s < splits; ++s) v += part[(size_t)s * stride + idx];
  return v;
}
__device__ __forceinline__ float4 sum_slabs4(const float* __restrict__ part, int splits, size_t stride, size_t idx) {
  if (splits <= kMaxSlabsMlp) {
    float4 t[kMaxSlabsMlp];
#pragma unroll
    for (int s = 0; s < kMaxSlabsMlp; ++s) t[s] = s < splits ? ld4(part + (size_t)s * stride + idx) : f4(0.f);
    float4 v = t[0];
#pragma unroll
    for (int s = 1; s < kMaxSlabsMlp; ++s) v = v + t[s];
    return v;
  }
  float4 v = ld4(part + idx);
  for (int s = 1; s < splits; ++s) v = v + ld4(part + (size_t)s * stride + idx);
  return v;
}
// Sum with four interleaved accumulators (slab s goes to accumulator s % 4, the tail beyond the last multiple of four
// to accumulator 0; result (a0 + a1) + (a2 + a3)), loads issued kChunk at a time: for the 36-slab logits GEMM, whose
// row kernels have one element per thread and nothing else to overlap the loads with -- one L2 round trip instead of nine.
#ifndef VLDD_NCE_CHUNK
#define VLDD_NCE_CHUNK 36
#endif
template <int kChunk = VLDD_NCE_CHUNK>
__device__ __forceinline__ float sum_slabs_ilp(const float* __restrict__ part, int splits, size_t stride, size_t idx) {
  static_assert(kChunk % 4 == 0, "chunks keep the slab -> accumulator mapping");
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  const int n4 = splits & ~3;
  for (int s0 = 0; s0 < splits; s0 += kChunk) {
    float t[kChunk];
#pragma unroll
    for (int u = 0; u < kChunk; ++u) t[u] = (s0 + u < splits) ? part[(size_t)(s0 + u) * stride + idx] : 0.f;
#pragma unroll
    for (int u = 0; u < kChunk; ++u) {
      if (s0 + u < n4) a[u & 3] += t[u];
      else a[0] += t[u];                        // tail slabs (and the zeros past `splits`)
    }
  }
  return (a[0] + a[1]) + (a[2] + a[3]);
}

// 128-bit path of the element-wise kernels: every pointer 16-byte aligned, row length and slab stride multiples of 4
__device__ __forceinline__ bool vec4_ok(int d, size_t stride, const void* a, const void* b = nullptr, const void* c = nullptr,
                                        const void* e = nullptr, const void* f = nullptr, const void* g = nullptr) {
  const uintptr_t bits = reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(c) |
                         reinterpret_cast<uintptr_t>(e) | reinterpret_cast<uintptr_t>(f) | reinterpret_cast<uintptr_t>(g);
  return (d & 3) == 0 && (stride & 3) == 0 && (bits & 15) == 0;
}

// ------------------------------------------------------------------------------------------------
// primal, forward
// ------------------------------------------------------------------------------------------------
// p = Y W1^T + b1 ; h = gelu(p)                                   [rows x d]
__global__ void __launch_bounds__(256) epi_p_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                    const float* __restrict__ b1, int rows, int d,
                                                    float* __restrict__ p, float* __restrict__ h) {
  pdl_enter();
  const size_t n = (size_t)rows * d;
  if (vec4_ok(d, stride, part, b1, p, h)) {
    for (size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * (size_t)gridDim.x * blockDim.x) {
      const float4 pv = sum_slabs4(part, splits, stride, i) + ld4(b1 + (int)(i % d));
      float4 hv;
      float d1, d2;
      gelu_parts(pv.x, hv.x, d1, d2);
      gelu_parts(pv.y, hv.y, d1, d2);
      gelu_parts(pv.z, hv.z, d1, d2);
      gelu_parts(pv.w, hv.w, d1, d2);
      st4(p + i, pv);
      st4(h + i, hv);
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const float pv = sum_slabs(part, splits, stride, i) + b1[c];
    float phi, d1, d2;
    gelu_parts(pv, phi, d1, d2);
    p[i] = pv;
    h[i] = phi;
  }
}

// r = mask*(h W2^T + b2) + p ; rhat = LN-normalised r ; z = gamma*rhat + beta ; yn = z/|z|        one CTA per row
// outputs: rhat (nullable), z (nullable), yn (nullable), rstd[row], nz[row] (nullable)
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                     const float* __restrict__ b2, const float* __restrict__ mask,
                                                     const float* __restrict__ p, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, int d,
                                                     float* __restrict__ rhat, float* __restrict__ z_out,
                                                     float* __restrict__ yn, float* __restrict__ rstd_out,
                                                     float* __restrict__ nz_out) {
  pdl_enter();
  extern __shared__ float rbuf[];  // d floats
  __shared__ float scratch[34];
  const int row = blockIdx.x;
  const size_t base = (size_t)row * d;
  float s = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float f = sum_slabs(part, splits, stride, base + j) + b2[j];
    if (mask) f *= mask[base + j];
    const float r = f + p[base + j];
    rbuf[j] = r;
    s += r;
  }
  const float mu = block_sum<float>(s, scratch) / d;
  float v = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float c = rbuf[j] - mu;
    v = fmaf(c, c, v);
  }
  const float rstd = rsqrtf(block_sum<float>(v, scratch) / d + kLnEps);
  float zz = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float rh = (rbuf[j] - mu) * rstd;
    const float z = fmaf(gamma[j], rh, beta[j]);
    if (rhat) rhat[base + j] = rh;
    rbuf[j] = z;
    zz = fmaf(z, z, zz);
  }
  const float nz = sqrtf(block_sum<float>(zz, scratch));
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float z = rbuf[j];
    if (z_out) z_out[base + j] = z;
    if (yn) yn[base + j] = z / nz;
  }
  if (threadIdx.x == 0) {
    if (rstd_out) rstd_out[row] = rstd;
    if (nz_out) nz_out[row] = nz;
  }
}

// x / |x| per row (distill.py:533 on the image-encoder output); also stores |x|
__global__ void __launch_bounds__(256) row_normalise_kernel(const float* __restrict__ x, int d, float* __restrict__ xn,
                                                            float* __restrict__ norm_out) {
  pdl_enter();
  __shared__ float scratch[34];
  const size_t base = (size_t)blockIdx.x * d;
  float s = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) s = fmaf(x[base + j], x[base + j], s);
  const float n = sqrtf(block_sum<float>(s, scratch));
  for (int j = threadIdx.x; j < d; j += blockDim.x) xn[base + j] = x[base + j] / n;
  if (threadIdx.x == 0 && norm_out) norm_out[blockIdx.x] = n;
}

// dU = c (dXn - xn <xn, dXn>) / |u|  with c = scale ? *scale : 1      backward of the row normalisation
__global__ void __launch_bounds__(256) row_normalise_bwd_kernel(const float* __restrict__ xn,
                                                                const float* __restrict__ norm,
                                                                const float* __restrict__ dxn,
                                                                const float* __restrict__ scale, int d,
                                                                float* __restrict__ dx) {
  pdl_enter();
  __shared__ float scratch[34];
  const size_t base = (size_t)blockIdx.x * d;
  float s = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) s = fmaf(xn[base + j], dxn[base + j], s);
  const float q = block_sum<float>(s, scratch);
  const float inv = (scale ? *scale : 1.0f) / norm[blockIdx.x];
  for (int j = threadIdx.x; j < d; j += blockDim.x) dx[base + j] = (dxn[base + j] - xn[base + j] * q) * inv;
}

// dst[b,:] = src[idx[b],:]
__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ src, const int64_t* __restrict__ idx,
                                                          int cols, float* __restrict__ dst) {
  pdl_enter();
  const size_t s = (size_t)idx[blockIdx.x] * cols, t = (size_t)blockIdx.x * cols;
  if ((cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    for (int j = threadIdx.x; j < (cols >> 2); j += blockDim.x)
      reinterpret_cast<float4*>(dst + t)[j] = reinterpret_cast<const float4*>(src + s)[j];
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) dst[t + j] = src[s + j];
  }
}

// All K steps at once: Yb[k][b,:] = Y[perms[k][b],:] and Xb[k][b,:] = Xn[perms[k][b],:]   (grid = (B, K, 2))
// An index outside [0, N) (the reference would raise IndexError, distill.py:512-513) is clamped so that no memory outside
// the operands is touched, and recorded in *bad_index: the call then reports NaN losses (poison_kernel) instead of silently
// training on the wrong rows.
__device__ __forceinline__ size_t checked_row(int64_t idx, int N, int* bad_index) {
  if (idx < 0 || idx >= N) {
    if (bad_index != nullptr) atomicExch(bad_index, 1);
    return idx < 0 ? 0 : (size_t)(N - 1);
  }
  return (size_t)idx;
}
// The caller's index array is reached through a one-pointer device slot (its address may change from call to call while the
// launch graph stays the same); the indices are copied into the workspace for the reverse sweep's scatter kernels.
__global__ void __launch_bounds__(256) gather_all_kernel(const float* __restrict__ Y, const float* __restrict__ Xn,
                                                         const int64_t* const* __restrict__ perms_slot,
                                                         int64_t* __restrict__ perms_copy, int B, int dt, int d,
                                                         float* __restrict__ Yb0, float* __restrict__ Xb0,
                                                         size_t step_stride, int N, int* __restrict__ bad_index) {
  pdl_enter();
  const int b = blockIdx.x, k = blockIdx.y;
  const bool isx = blockIdx.z == 1;
  const int cols = isx ? d : dt;
  const int64_t index = (*perms_slot)[(size_t)k * B + b];
  if (!isx && threadIdx.x == 0) perms_copy[(size_t)k * B + b] = index;
  const size_t row = checked_row(index, N, bad_index);
  const float* src = (isx ? Xn : Y) + row * cols;
  float* dst = (isx ? Xb0 : Yb0) + (size_t)k * step_stride + (size_t)b * cols;
  if ((cols & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    for (int j = threadIdx.x; j < (cols >> 2); j += blockDim.x)
      reinterpret_cast<float4*>(dst)[j] = reinterpret_cast<const float4*>(src)[j];
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) dst[j] = src[j];
  }
}

// dst[idx[b],:] += coef * sum_slabs(part)[b,:]   with coef = -(*lr) * (scale ? *scale : 1)
// The reference's minibatch indices are unique within a step (randperm, distill.py:510-511), but the C ABI takes arbitrary
// caller permutations: the accumulation is a 128-bit reduction at the L2 (red.global.add.v4.f32), so duplicate rows add up
// correctly instead of racing.  With unique indices every element receives one add per step: bit-reproducible as before.
__device__ __forceinline__ void red_add4(float* p, float4 v) {
  asm volatile("red.global.v4.f32.add [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                               const int64_t* __restrict__ idx, int cols,
                                                               const float* __restrict__ lr,
                                                               const float* __restrict__ scale,
                                                               float* __restrict__ dst, int N) {
  pdl_enter();
  const float coef = -(*lr) * (scale ? *scale : 1.0f);
  const size_t s = (size_t)blockIdx.x * cols, t = checked_row(idx[blockIdx.x], N, nullptr) * cols;
  if ((cols & 3) == 0 && (stride & 3) == 0 &&
      ((reinterpret_cast<uintptr_t>(part) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
    for (int j = threadIdx.x; j < (cols >> 2); j += blockDim.x) {
      const float4 v = sum_slabs4(part, splits, stride, s + 4 * j);
      red_add4(dst + t + 4 * j, make_float4(coef * v.x, coef * v.y, coef * v.z, coef * v.w));
    }
  } else {
    for (int j = threadIdx.x; j < cols; j += blockDim.x) atomicAdd(dst + t + j, coef * sum_slabs(part, splits, stride, s + j));
  }
}

// Last kernel of the call: out5[0..2] = {num, den, num / den} from the fp64 block partials the matching-loss pass left
// (added in index order: deterministic), then out5[0..4] := NaN when a minibatch index was out of range (see checked_row).
__global__ void __launch_bounds__(256) finalize_kernel(const double* __restrict__ parts, int n_parts, const float* __restrict__ den_p,
                                                       const int* __restrict__ bad_index, float* __restrict__ out5) {
  pdl_enter();
  __shared__ double scratch[34];
  double s = 0.0;
  for (int i = threadIdx.x; i < n_parts; i += blockDim.x) s += parts[i];
  s = block_sum<double>(s, scratch);
  const bool bad = bad_index != nullptr && *bad_index != 0;
  if (threadIdx.x == 0) {
    const float den = *den_p;
    out5[0] = (float)s;
    out5[1] = den;
    out5[2] = (float)(s / (double)den);
  }
  __syncthreads();
  if (threadIdx.x < 5 && bad) out5[threadIdx.x] = __int_as_float(0x7fc00000);
}

// ------------------------------------------------------------------------------------------------
// InfoNCE, primal.  S = scale * Xn Yn^T  (B x B, small, L2-resident)
// ------------------------------------------------------------------------------------------------
// row i: S[i,:] = scale * sum_slabs ; lse_r[i]
__global__ void __launch_bounds__(128) nce_rows_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                       const float* __restrict__ scale, int B, int ld,
                                                       float* __restrict__ S, float* __restrict__ lse_r) {
  pdl_enter();
  __shared__ float scratch[34];
  const int i = blockIdx.x;
  const float sc = *scale;
  float mx = -INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float v = sc * sum_slabs_ilp(part, splits, stride, (size_t)i * B + j);
    S[(size_t)i * ld + j] = v;
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, scratch);
  float se = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) se += expf(S[(size_t)i * ld + j] - mx);
  se = block_sum<float>(se, scratch);
  if (threadIdx.x == 0) lse_r[i] = mx + logf(se);
}
// column j: lse_c[j]
__global__ void __launch_bounds__(128) nce_cols_kernel(const float* __restrict__ S, int B, int ld,
                                                       float* __restrict__ lse_c) {
  pdl_enter();
  __shared__ float scratch[34];
  const int j = blockIdx.x;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < B; i += blockDim.x) mx = fmaxf(mx, S[(size_t)i * ld + j]);
  mx = block_max(mx, scratch);
  float se = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) se += expf(S[(size_t)i * ld + j] - mx);
  se = block_sum<float>(se, scratch);
  if (threadIdx.x == 0) lse_c[j] = mx + logf(se);
}
// G = (softmax_rows + softmax_cols - 2I) / (2B); block 0 also writes the loss
__global__ void __launch_bounds__(128) nce_grad_kernel(const float* __restrict__ S, const float* __restrict__ lse_r,
                                                       const float* __restrict__ lse_c, int B, int ld,
                                                       float* __restrict__ G, float* __restrict__ loss_out) {
  pdl_enter();
  __shared__ float scratch[34];
  const int i = blockIdx.x;
  const float inv2B = 0.5f / B, lr_i = lse_r[i];
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const float s = S[(size_t)i * ld + j];
    float g = expf(s - lr_i) + expf(s - lse_c[j]);
    if (j == i) g -= 2.0f;
    G[(size_t)i * ld + j] = g * inv2B;
  }
  if (blockIdx.x == 0) {
    float acc = 0.f;
    for (int r = threadIdx.x; r < B; r += blockDim.x) acc += (lse_r[r] - S[(size_t)r * ld + r]) + (lse_c[r] - S[(size_t)r * ld + r]);
    acc = block_sum<float>(acc, scratch);
    if (threadIdx.x == 0 && loss_out) *loss_out = acc * inv2B;
  }
}

// ------------------------------------------------------------------------------------------------
// primal, backward
// ------------------------------------------------------------------------------------------------
// dyn = scale * raw ; q = <yn,dyn> ; dz = (dyn - yn q)/nz ; drhat = gamma dz ; dr = rstd (drhat - m1 - rhat m2) ; df = mask dr
__global__ void __launch_bounds__(256) norm_ln_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                                                          const float* __restrict__ yn, const float* __restrict__ nz_p,
                                                          const float* __restrict__ rhat, const float* __restrict__ rstd_p,
                                                          const float* __restrict__ gamma, const float* __restrict__ mask,
                                                          int d, float* __restrict__ dyn, float* __restrict__ q_out,
                                                          float* __restrict__ dz, float* __restrict__ dr,
                                                          float* __restrict__ df) {
  pdl_enter();
  __shared__ float scratch[34];
  const int row = blockIdx.x;
  const size_t base = (size_t)row * d;
  const float sc = *scale, nz = nz_p[row], rstd = rstd_p[row];
  float s = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float v = sc * raw[base + j];
    dyn[base + j] = v;
    s = fmaf(yn[base + j], v, s);
  }
  const float q = block_sum<float>(s, scratch);
  float s1 = 0.f, s2 = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float v = (dyn[base + j] - yn[base + j] * q) / nz;
    dz[base + j] = v;
    const float drh = gamma[j] * v;
    s1 += drh;
    s2 = fmaf(drh, rhat[base + j], s2);
  }
  const float m1 = block_sum<float>(s1, scratch) / d;
  const float m2 = block_sum<float>(s2, scratch) / d;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float drh = gamma[j] * dz[base + j];
    const float v = rstd * (drh - m1 - rhat[base + j] * m2);
    dr[base + j] = v;
    df[base + j] = mask ? v * mask[base + j] : v;
  }
  if (threadIdx.x == 0) q_out[row] = q;
}

// dh = df W2 ; dp = dh * gelu'(p) + dr
__global__ void __launch_bounds__(256) epi_dp_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                     const float* __restrict__ p, const float* __restrict__ dr,
                                                     size_t n, float* __restrict__ dh, float* __restrict__ dp) {
  pdl_enter();
  if (vec4_ok((int)(n & 3), stride, part, p, dr, dh, dp)) {
    for (size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * (size_t)gridDim.x * blockDim.x) {
      const float4 v = sum_slabs4(part, splits, stride, i), pv = ld4(p + i), drv = ld4(dr + i);
      float phi, d2;
      float4 d1;
      gelu_parts(pv.x, phi, d1.x, d2);
      gelu_parts(pv.y, phi, d1.y, d2);
      gelu_parts(pv.z, phi, d1.z, d2);
      gelu_parts(pv.w, phi, d1.w, d2);
      st4(dh + i, v);
      st4(dp + i, make_float4(fmaf(v.x, d1.x, drv.x), fmaf(v.y, d1.y, drv.y), fmaf(v.z, d1.z, drv.z), fmaf(v.w, d1.w, drv.w)));
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = sum_slabs(part, splits, stride, i);
    float phi, d1, d2;
    gelu_parts(p[i], phi, d1, d2);
    dh[i] = v;
    dp[i] = fmaf(v, d1, dr[i]);
  }
}

// Column sums over the batch and the fused update of the small parameters:
//   db1 = sum dp ; db2 = sum df ; dgamma = sum dz*rhat ; dbeta = sum dz ;  dst = src - lr * grad   (src nullable = 0)
// block = 32 columns x 8 row groups (each warp reads 128 contiguous bytes of one row); fixed-order smem reduction.
__global__ void __launch_bounds__(256) colsum_update_kernel(const float* __restrict__ dp, const float* __restrict__ df,
                                                            const float* __restrict__ dz, const float* __restrict__ rhat,
                                                            int B, int d, const float* __restrict__ lr,
                                                            const float* __restrict__ src_b1, float* __restrict__ dst_b1,
                                                            const float* __restrict__ src_b2, float* __restrict__ dst_b2,
                                                            const float* __restrict__ src_g, float* __restrict__ dst_g,
                                                            const float* __restrict__ src_b, float* __restrict__ dst_b) {
  pdl_enter();
  __shared__ float red[4][16][17];
  const int cx = threadIdx.x & 15, rg = threadIdx.x >> 4;
  const int n = blockIdx.x * 16 + cx;
  float a1 = 0.f, a2 = 0.f, ag = 0.f, ab = 0.f;
  if (n < d) {
    for (int b = rg; b < B; b += 16) {
      const size_t i = (size_t)b * d + n;
      a1 += dp[i];
      a2 += df[i];
      const float z = dz[i];
      ag = fmaf(z, rhat[i], ag);
      ab += z;
    }
  }
  red[0][rg][cx] = a1; red[1][rg][cx] = a2; red[2][rg][cx] = ag; red[3][rg][cx] = ab;
  __syncthreads();
  if (rg < 4 && n < d) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) t += red[rg][r][cx];
    const float l = *lr;
    const float* src = rg == 0 ? src_b1 : rg == 1 ? src_b2 : rg == 2 ? src_g : src_b;
    float* dst = rg == 0 ? dst_b1 : rg == 1 ? dst_b2 : rg == 2 ? dst_g : dst_b;
    dst[n] = (src ? src[n] : 0.f) - l * t;
  }
}

// ------------------------------------------------------------------------------------------------
// tangent, forward
// ------------------------------------------------------------------------------------------------
// pd = Y V1^T + c1 ; hd = gelu'(p) pd
__global__ void __launch_bounds__(256) epi_pd_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                     const float* __restrict__ c1, const float* __restrict__ p, int rows,
                                                     int d, float* __restrict__ pd, float* __restrict__ hd) {
  pdl_enter();
  const size_t n = (size_t)rows * d;
  if (vec4_ok(d, stride, part, c1, p, pd, hd)) {
    for (size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * (size_t)gridDim.x * blockDim.x) {
      const float4 v = sum_slabs4(part, splits, stride, i) + ld4(c1 + (int)(i % d)), pv = ld4(p + i);
      float phi, d2;
      float4 d1;
      gelu_parts(pv.x, phi, d1.x, d2);
      gelu_parts(pv.y, phi, d1.y, d2);
      gelu_parts(pv.z, phi, d1.z, d2);
      gelu_parts(pv.w, phi, d1.w, d2);
      st4(pd + i, v);
      st4(hd + i, d1 * v);
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d);
    const float v = sum_slabs(part, splits, stride, i) + c1[c];
    float phi, d1, d2;
    gelu_parts(p[i], phi, d1, d2);
    pd[i] = v;
    hd[i] = d1 * v;
  }
}

// fd = (hd W2^T + h V2^T) + c2 ; rd = mask fd + pd ; t = mean(rhat rd) ; rhatd = rstd (rd - mean rd - rhat t)
// zd = gammad rhat + gamma rhatd + betad ; nzd = <yn, zd> ; ynd = (zd - yn nzd)/nz
__global__ void __launch_bounds__(256) ln_tangent_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                         const float* __restrict__ c2, const float* __restrict__ mask,
                                                         const float* __restrict__ pd, const float* __restrict__ rhat,
                                                         const float* __restrict__ rstd_p, const float* __restrict__ yn,
                                                         const float* __restrict__ nz_p, const float* __restrict__ gamma,
                                                         const float* __restrict__ gammad, const float* __restrict__ betad,
                                                         int d, float* __restrict__ rhatd, float* __restrict__ ynd,
                                                         float* __restrict__ t_out, float* __restrict__ nzd_out) {
  pdl_enter();
  extern __shared__ float buf[];  // d floats
  __shared__ float scratch[34];
  const int row = blockIdx.x;
  const size_t base = (size_t)row * d;
  const float rstd = rstd_p[row], nz = nz_p[row];
  float s1 = 0.f, s2 = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float f = sum_slabs(part, splits, stride, base + j) + c2[j];
    if (mask) f *= mask[base + j];
    const float rd = f + pd[base + j];
    buf[j] = rd;
    s1 += rd;
    s2 = fmaf(rhat[base + j], rd, s2);
  }
  const float mrd = block_sum<float>(s1, scratch) / d;
  const float t = block_sum<float>(s2, scratch) / d;
  float s3 = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float rh = rhat[base + j];
    const float rhd = rstd * (buf[j] - mrd - rh * t);
    rhatd[base + j] = rhd;
    const float zd = gammad[j] * rh + gamma[j] * rhd + betad[j];
    buf[j] = zd;
    s3 = fmaf(yn[base + j], zd, s3);
  }
  const float nzd = block_sum<float>(s3, scratch);
  for (int j = threadIdx.x; j < d; j += blockDim.x) ynd[base + j] = (buf[j] - yn[base + j] * nzd) / nz;
  if (threadIdx.x == 0) { t_out[row] = t; nzd_out[row] = nzd; }
}

// ------------------------------------------------------------------------------------------------
// InfoNCE, tangent
// ------------------------------------------------------------------------------------------------
// row i: Sd[i,:] = scale * slabs ; rho_i = sum_j Pr_ij Sd_ij ; rowLd[i] = sum_j G_ij Sd_ij
// (scale_dot != nullptr: the logit scale itself moves along the direction, Sd += (*scale_dot / scale) * S)
__global__ void __launch_bounds__(128) nce_t_rows_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                         const float* __restrict__ scale, const float* __restrict__ S,
                                                         const float* __restrict__ lse_r, const float* __restrict__ G,
                                                         int B, int ld, float* __restrict__ Sd, float* __restrict__ rho,
                                                         float* __restrict__ rowLd,
                                                         const float* __restrict__ scale_dot = nullptr) {
  pdl_enter();
  __shared__ float scratch[34];
  const int i = blockIdx.x;
  const float sc = *scale, l = lse_r[i];
  const float rel = scale_dot ? (*scale_dot) / sc : 0.f;
  float a = 0.f, b = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const size_t ij = (size_t)i * ld + j;
    float v = sc * sum_slabs_ilp(part, splits, stride, (size_t)i * B + j);
    if (scale_dot) v = fmaf(rel, S[ij], v);
    Sd[ij] = v;
    a = fmaf(expf(S[ij] - l), v, a);
    b = fmaf(G[ij], v, b);
  }
  a = block_sum<float>(a, scratch);
  b = block_sum<float>(b, scratch);
  if (threadIdx.x == 0) { rho[i] = a; rowLd[i] = b; }
}
// column j: kap_j = sum_i Pc_ij Sd_ij
__global__ void __launch_bounds__(128) nce_t_cols_kernel(const float* __restrict__ S, const float* __restrict__ lse_c,
                                                         const float* __restrict__ Sd, int B, int ld,
                                                         float* __restrict__ kap) {
  pdl_enter();
  __shared__ float scratch[34];
  const int j = blockIdx.x;
  const float l = lse_c[j];
  float a = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const size_t ij = (size_t)i * ld + j;
    a = fmaf(expf(S[ij] - l), Sd[ij], a);
  }
  a = block_sum<float>(a, scratch);
  if (threadIdx.x == 0) kap[j] = a;
}
// Gd = (Pr (Sd - rho_i) + Pc (Sd - kap_j)) / (2B) ; rowGdS[i] = sum_j Gd_ij S_ij
__global__ void __launch_bounds__(128) nce_t_grad_kernel(const float* __restrict__ S, const float* __restrict__ lse_r,
                                                         const float* __restrict__ lse_c, const float* __restrict__ Sd,
                                                         const float* __restrict__ rho, const float* __restrict__ kap,
                                                         int B, int ld, float* __restrict__ Gd,
                                                         float* __restrict__ rowGdS) {
  pdl_enter();
  __shared__ float scratch[34];
  const int i = blockIdx.x;
  const float inv2B = 0.5f / B, l = lse_r[i], rh = rho[i];
  float a = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    const size_t ij = (size_t)i * ld + j;
    const float s = S[ij], sd = Sd[ij];
    const float g = (expf(s - l) * (sd - rh) + expf(s - lse_c[j]) * (sd - kap[j])) * inv2B;
    Gd[ij] = g;
    a = fmaf(g, s, a);
  }
  a = block_sum<float>(a, scratch);
  if (threadIdx.x == 0) rowGdS[i] = a;
}
// one block: Ld = sum rowLd ; dlr -= Ld ; dscale -= lr * (sum rowGdS + Ld) / scale
__global__ void __launch_bounds__(128) nce_t_finish_kernel(const float* __restrict__ rowLd,
                                                           const float* __restrict__ rowGdS, int B,
                                                           const float* __restrict__ lr, const float* __restrict__ scale,
                                                           float* __restrict__ dlr, float* __restrict__ dscale) {
  pdl_enter();
  __shared__ float scratch[34];
  float a = 0.f, b = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) { a += rowLd[i]; b += rowGdS[i]; }
  a = block_sum<float>(a, scratch);
  b = block_sum<float>(b, scratch);
  if (threadIdx.x == 0) {
    *dlr -= a;
    *dscale -= (*lr) * (b + a) / (*scale);
  }
}

// ------------------------------------------------------------------------------------------------
// tangent, backward
// ------------------------------------------------------------------------------------------------
// dynd = scale*raw ; qd = <ynd,dyn> + <yn,dynd> ; dzd = (dynd - ynd q - yn qd)/nz - dz nzd/nz
// drhatd = gammad dz + gamma dzd ; drd = -(rstd t) dr + rstd (drhatd - m1d - rhatd m2 - rhat m2d) ; dfd = mask drd
__global__ void __launch_bounds__(256) norm_ln_bwd_tangent_kernel(
    const float* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ yn,
    const float* __restrict__ ynd, const float* __restrict__ dyn, const float* __restrict__ q_p,
    const float* __restrict__ nz_p, const float* __restrict__ nzd_p, const float* __restrict__ dz,
    const float* __restrict__ rhat, const float* __restrict__ rhatd, const float* __restrict__ rstd_p,
    const float* __restrict__ t_p, const float* __restrict__ dr, const float* __restrict__ gamma,
    const float* __restrict__ gammad, const float* __restrict__ mask, int d, float* __restrict__ dzd,
    float* __restrict__ drd, float* __restrict__ dfd) {
  pdl_enter();
  extern __shared__ float buf[];  // d floats (dynd, then drhatd)
  __shared__ float scratch[34];
  const int row = blockIdx.x;
  const size_t base = (size_t)row * d;
  const float sc = *scale, nz = nz_p[row], nzd = nzd_p[row], q = q_p[row], rstd = rstd_p[row], t = t_p[row];
  float s = 0.f;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float v = sc * raw[base + j];
    buf[j] = v;
    s = fmaf(ynd[base + j], dyn[base + j], s);
    s = fmaf(yn[base + j], v, s);
  }
  const float qd = block_sum<float>(s, scratch);
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  const float inz = 1.0f / nz;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float dzv = dz[base + j];
    const float v = (buf[j] - ynd[base + j] * q - yn[base + j] * qd) * inz - dzv * (nzd * inz);
    dzd[base + j] = v;
    const float drh = gamma[j] * dzv;
    const float drhd = gammad[j] * dzv + gamma[j] * v;
    buf[j] = drhd;
    s1 += drhd;
    s2 = fmaf(drhd, rhat[base + j], s2);
    s2 = fmaf(drh, rhatd[base + j], s2);
    s3 = fmaf(drh, rhat[base + j], s3);
  }
  const float m1d = block_sum<float>(s1, scratch) / d;
  const float m2d = block_sum<float>(s2, scratch) / d;
  const float m2 = block_sum<float>(s3, scratch) / d;
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    const float v = -(rstd * t) * dr[base + j] + rstd * (buf[j] - m1d - rhatd[base + j] * m2 - rhat[base + j] * m2d);
    drd[base + j] = v;
    dfd[base + j] = mask ? v * mask[base + j] : v;
  }
}

// dhd = dfd W2 + df V2 ; dpd = dhd gelu'(p) + dh gelu''(p) pd + drd
__global__ void __launch_bounds__(256) epi_dpd_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                      const float* __restrict__ p, const float* __restrict__ pd,
                                                      const float* __restrict__ dh, const float* __restrict__ drd,
                                                      size_t n, float* __restrict__ dpd) {
  pdl_enter();
  if (vec4_ok((int)(n & 3), stride, part, p, pd, dh, drd, dpd)) {
    for (size_t i = 4 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); i < n; i += 4 * (size_t)gridDim.x * blockDim.x) {
      const float4 v = sum_slabs4(part, splits, stride, i), pv = ld4(p + i), pdv = ld4(pd + i), dhv = ld4(dh + i),
                   drv = ld4(drd + i);
      float phi;
      float4 d1, d2;
      gelu_parts(pv.x, phi, d1.x, d2.x);
      gelu_parts(pv.y, phi, d1.y, d2.y);
      gelu_parts(pv.z, phi, d1.z, d2.z);
      gelu_parts(pv.w, phi, d1.w, d2.w);
      st4(dpd + i, make_float4(fmaf(v.x, d1.x, fmaf(dhv.x * d2.x, pdv.x, drv.x)), fmaf(v.y, d1.y, fmaf(dhv.y * d2.y, pdv.y, drv.y)),
                               fmaf(v.z, d1.z, fmaf(dhv.z * d2.z, pdv.z, drv.z)), fmaf(v.w, d1.w, fmaf(dhv.w * d2.w, pdv.w, drv.w))));
    }
    return;
  }
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = sum_slabs(part, splits, stride, i);
    float phi, d1, d2;
    gelu_parts(p[i], phi, d1, d2);
    dpd[i] = fmaf(v, d1, fmaf(dh[i] * d2, pd[i], drd[i]));
  }
}

// db1d = sum dpd ; db2d = sum dfd ; dgammad = sum (dzd rhat + dz rhatd) ; dbetad = sum dzd ;  dst = src - lr * (.)
__global__ void __launch_bounds__(256) colsum_tangent_update_kernel(
    const float* __restrict__ dpd, const float* __restrict__ dfd, const float* __restrict__ dzd,
    const float* __restrict__ dz, const float* __restrict__ rhat, const float* __restrict__ rhatd, int B, int d,
    const float* __restrict__ lr, const float* __restrict__ src_b1, float* __restrict__ dst_b1,
    const float* __restrict__ src_b2, float* __restrict__ dst_b2, const float* __restrict__ src_g,
    float* __restrict__ dst_g, const float* __restrict__ src_b, float* __restrict__ dst_b) {
  pdl_enter();
  __shared__ float red[4][16][17];
  const int cx = threadIdx.x & 15, rg = threadIdx.x >> 4;
  const int n = blockIdx.x * 16 + cx;
  float a1 = 0.f, a2 = 0.f, ag = 0.f, ab = 0.f;
  if (n < d) {
    for (int b = rg; b < B; b += 16) {
      const size_t i = (size_t)b * d + n;
      a1 += dpd[i];
      a2 += dfd[i];
      const float zd = dzd[i];
      ag = fmaf(zd, rhat[i], ag);
      ag = fmaf(dz[i], rhatd[i], ag);
      ab += zd;
    }
  }
  red[0][rg][cx] = a1; red[1][rg][cx] = a2; red[2][rg][cx] = ag; red[3][rg][cx] = ab;
  __syncthreads();
  if (rg < 4 && n < d) {
    float t = 0.f;
#pragma unroll
    for (int r = 0; r < 16; ++r) t += red[rg][r][cx];
    const float l = *lr;
    const float* src = rg == 0 ? src_b1 : rg == 1 ? src_b2 : rg == 2 ? src_g : src_b;
    float* dst = rg == 0 ? dst_b1 : rg == 1 ? dst_b2 : rg == 2 ? dst_g : dst_b;
    dst[n] = src[n] - l * t;
  }
}

}  // namespace vldd
