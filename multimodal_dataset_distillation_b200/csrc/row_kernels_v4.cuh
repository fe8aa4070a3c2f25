// Register-resident float4 versions of the four LayerNorm/normalise row kernels (one CTA of 256 threads per row,
// d % 4 == 0, d <= 1024 * NV).  Same math as the generic kernels in head_kernels.cuh (which remain the fallback for
// other d); each row is read from global memory once and kept in registers across the dependent reductions.
#pragma once
#include "head_kernels.cuh"

namespace vldd {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float4 f4(float a) { return make_float4(a, a, a, a); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator-(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float4 b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
__device__ __forceinline__ float4 operator*(float a, float4 b) { return make_float4(a * b.x, a * b.y, a * b.z, a * b.w); }
__device__ __forceinline__ float hsum(float4 a) { return (a.x + a.y) + (a.z + a.w); }
__device__ __forceinline__ float hdot(float4 a, float4 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w))); }

__device__ __forceinline__ float4 sum_slabs4(const float* __restrict__ part, int splits, size_t stride, size_t idx) {
  float4 v = ld4(part + idx);
  for (int s = 1; s < splits; ++s) v = v + ld4(part + (size_t)s * stride + idx);
  return v;
}

// two simultaneous block sums (one barrier round trip instead of two)
__device__ __forceinline__ float2 block_sum2(float a, float b, float2* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0) scratch[wid] = make_float2(a, b);
  __syncthreads();
  float2 t = lane < nw ? scratch[lane] : make_float2(0.f, 0.f);
  t.x = warp_sum(t.x);
  t.y = warp_sum(t.y);
  return t;
}
__device__ __forceinline__ float block_sum1(float a, float2* scratch) { return block_sum2(a, 0.f, scratch).x; }

template <int NV>
__global__ void __launch_bounds__(256) ln_fwd_v4_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                        const float* __restrict__ b2, const float* __restrict__ mask,
                                                        const float* __restrict__ p, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, int d, float* __restrict__ rhat,
                                                        float* __restrict__ z_out, float* __restrict__ yn,
                                                        float* __restrict__ rstd_out, float* __restrict__ nz_out) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  float4 r[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    r[i] = f4(0.f);
    if (j < d4) {
      float4 f = sum_slabs4(part, splits, stride, base + 4 * j) + ld4(b2 + 4 * j);
      if (mask) f = f * ld4(mask + base + 4 * j);
      r[i] = f + ld4(p + base + 4 * j);
      s += hsum(r[i]);
    }
  }
  const float mu = block_sum1(s, scratch) / d;
  float v = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) { r[i] = r[i] - f4(mu); v += hdot(r[i], r[i]); }
  }
  const float rstd = rsqrtf(block_sum1(v, scratch) / d + kLnEps);
  float zz = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 rh = rstd * r[i];
      if (rhat) st4(rhat + base + 4 * j, rh);
      r[i] = ld4(gamma + 4 * j) * rh + ld4(beta + 4 * j);
      zz += hdot(r[i], r[i]);
    }
  }
  const float nz = sqrtf(block_sum1(zz, scratch));
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      if (z_out) st4(z_out + base + 4 * j, r[i]);
      if (yn) st4(yn + base + 4 * j, make_float4(r[i].x / nz, r[i].y / nz, r[i].z / nz, r[i].w / nz));
    }
  }
  if (threadIdx.x == 0) {
    if (rstd_out) rstd_out[row] = rstd;
    if (nz_out) nz_out[row] = nz;
  }
}

template <int NV>
__global__ void __launch_bounds__(256) norm_ln_bwd_v4_kernel(const float* __restrict__ raw, const float* __restrict__ scale,
                                                             const float* __restrict__ yn, const float* __restrict__ nz_p,
                                                             const float* __restrict__ rhat, const float* __restrict__ rstd_p,
                                                             const float* __restrict__ gamma, const float* __restrict__ mask,
                                                             int d, float* __restrict__ dyn, float* __restrict__ q_out,
                                                             float* __restrict__ dz, float* __restrict__ dr,
                                                             float* __restrict__ df) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  const float sc = *scale, inz = 1.0f / nz_p[row], rstd = rstd_p[row];
  float4 a[NV], y[NV], rh[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      a[i] = sc * ld4(raw + base + 4 * j);
      y[i] = ld4(yn + base + 4 * j);
      rh[i] = ld4(rhat + base + 4 * j);
      st4(dyn + base + 4 * j, a[i]);
      s += hdot(y[i], a[i]);
    }
  }
  const float q = block_sum1(s, scratch);
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 dzv = inz * (a[i] - q * y[i]);
      st4(dz + base + 4 * j, dzv);
      a[i] = ld4(gamma + 4 * j) * dzv;     // drhat
      s1 += hsum(a[i]);
      s2 += hdot(a[i], rh[i]);
    }
  }
  const float2 mm = block_sum2(s1, s2, scratch);
  const float m1 = mm.x / d, m2 = mm.y / d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 v = rstd * (a[i] - f4(m1) - m2 * rh[i]);
      st4(dr + base + 4 * j, v);
      st4(df + base + 4 * j, mask ? v * ld4(mask + base + 4 * j) : v);
    }
  }
  if (threadIdx.x == 0) q_out[row] = q;
}

template <int NV>
__global__ void __launch_bounds__(256) ln_tangent_v4_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                            const float* __restrict__ c2, const float* __restrict__ mask,
                                                            const float* __restrict__ pd, const float* __restrict__ rhat,
                                                            const float* __restrict__ rstd_p, const float* __restrict__ yn,
                                                            const float* __restrict__ nz_p, const float* __restrict__ gamma,
                                                            const float* __restrict__ gammad, const float* __restrict__ betad,
                                                            int d, float* __restrict__ rhatd, float* __restrict__ ynd,
                                                            float* __restrict__ t_out, float* __restrict__ nzd_out) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  const float rstd = rstd_p[row], nz = nz_p[row];
  float4 rd[NV], rh[NV], y[NV];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      float4 f = sum_slabs4(part, splits, stride, base + 4 * j) + ld4(c2 + 4 * j);
      if (mask) f = f * ld4(mask + base + 4 * j);
      rd[i] = f + ld4(pd + base + 4 * j);
      rh[i] = ld4(rhat + base + 4 * j);
      y[i] = ld4(yn + base + 4 * j);
      s1 += hsum(rd[i]);
      s2 += hdot(rh[i], rd[i]);
    }
  }
  const float2 mm = block_sum2(s1, s2, scratch);
  const float mrd = mm.x / d, t = mm.y / d;
  float s3 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 rhd = rstd * (rd[i] - f4(mrd) - t * rh[i]);
      st4(rhatd + base + 4 * j, rhd);
      rd[i] = ld4(gammad + 4 * j) * rh[i] + ld4(gamma + 4 * j) * rhd + ld4(betad + 4 * j);   // zd
      s3 += hdot(y[i], rd[i]);
    }
  }
  const float nzd = block_sum1(s3, scratch);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 o = rd[i] - nzd * y[i];
      st4(ynd + base + 4 * j, make_float4(o.x / nz, o.y / nz, o.z / nz, o.w / nz));
    }
  }
  if (threadIdx.x == 0) { t_out[row] = t; nzd_out[row] = nzd; }
}

template <int NV>
__global__ void __launch_bounds__(256) norm_ln_bwd_tangent_v4_kernel(
    const float* __restrict__ raw, const float* __restrict__ scale, const float* __restrict__ yn,
    const float* __restrict__ ynd, const float* __restrict__ dyn, const float* __restrict__ q_p,
    const float* __restrict__ nz_p, const float* __restrict__ nzd_p, const float* __restrict__ dz,
    const float* __restrict__ rhat, const float* __restrict__ rhatd, const float* __restrict__ rstd_p,
    const float* __restrict__ t_p, const float* __restrict__ dr, const float* __restrict__ gamma,
    const float* __restrict__ gammad, const float* __restrict__ mask, int d, float* __restrict__ dzd,
    float* __restrict__ drd, float* __restrict__ dfd) {
  pdl_enter();
  __shared__ float2 scratch[32];
  const int row = blockIdx.x, d4 = d >> 2;
  const size_t base = (size_t)row * d;
  const float sc = *scale, nz = nz_p[row], nzd = nzd_p[row], q = q_p[row], rstd = rstd_p[row], t = t_p[row];
  const float inz = 1.0f / nz;
  float4 a[NV], y[NV], yd[NV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      a[i] = sc * ld4(raw + base + 4 * j);           // dynd
      y[i] = ld4(yn + base + 4 * j);
      yd[i] = ld4(ynd + base + 4 * j);
      s += hdot(yd[i], ld4(dyn + base + 4 * j)) + hdot(y[i], a[i]);
    }
  }
  const float qd = block_sum1(s, scratch);
  float4 rh[NV], rhd[NV];
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 dzv = ld4(dz + base + 4 * j);
      const float4 v = inz * (a[i] - q * yd[i] - qd * y[i]) - (nzd * inz) * dzv;
      st4(dzd + base + 4 * j, v);
      const float4 g = ld4(gamma + 4 * j);
      const float4 drh = g * dzv;
      a[i] = ld4(gammad + 4 * j) * dzv + g * v;      // drhatd
      rh[i] = ld4(rhat + base + 4 * j);
      rhd[i] = ld4(rhatd + base + 4 * j);
      s1 += hsum(a[i]);
      s2 += hdot(a[i], rh[i]) + hdot(drh, rhd[i]);
      s3 += hdot(drh, rh[i]);
    }
  }
  const float2 mm = block_sum2(s1, s2, scratch);
  const float m1d = mm.x / d, m2d = mm.y / d;
  const float m2 = block_sum1(s3, scratch) / d;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int j = threadIdx.x + 256 * i;
    if (j < d4) {
      const float4 v = (-(rstd * t)) * ld4(dr + base + 4 * j) + rstd * (a[i] - f4(m1d) - m2 * rhd[i] - m2d * rh[i]);
      st4(drd + base + 4 * j, v);
      st4(dfd + base + 4 * j, mask ? v * ld4(mask + base + 4 * j) : v);
    }
  }
}

inline bool row_v4_ok(int d) { return d % 4 == 0 && d <= 3072; }
#define VLDD_UNPACK(...) __VA_ARGS__
#define VLDD_ROW_V4_DISPATCH(d, KERNEL, CFG, ARGS)                                         \
  do {                                                                                     \
    if ((d) <= 1024) launch_k(KERNEL<1>, VLDD_UNPACK CFG, VLDD_UNPACK ARGS);               \
    else if ((d) <= 2048) launch_k(KERNEL<2>, VLDD_UNPACK CFG, VLDD_UNPACK ARGS);          \
    else launch_k(KERNEL<3>, VLDD_UNPACK CFG, VLDD_UNPACK ARGS);                           \
  } while (0)

}  // namespace vldd
